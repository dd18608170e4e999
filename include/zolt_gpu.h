/* zolt_gpu.h -- C ABI of the B200-native prover hot path for Zolt (MatteoMer/zolt).
 *
 * The reference has NO FFI today; these entry points are what a `zig/zolt_gpu.zig`
 * shim binds with @cImport so that the existing Zig call signatures stay intact.  Each function
 * cites the reference interface it replaces (paths relative to /root/reference).
 *
 * Conventions
 *  - Every function returns an int32 status (ZOLT_OK == 0).  Nothing throws, nothing calls back.
 *    Calls block until their outputs are valid unless the name ends in `_async`.
 *  - There is NO CPU fallback: with no usable CUDA device zolt_gpu_init fails with
 *    ZOLT_ERR_NOGPU and every other call needs a context.
 *  - Field elements (zolt_fr / zolt_fq): uint64_t[4], little-endian limbs, Montgomery form
 *    a*2^256 mod p, canonical (< p) -- bit-identical to `BN254Scalar.limbs` /
 *    `MontgomeryField.limbs` (src/field/mod.zig:583-585, :124-130).
 *  - G1 affine points: uint64_t xy[8] = x limbs then y limbs (Fq, Montgomery) plus a separate
 *    uint8_t infinity flag.  Identity is x=0, y=0, infinity=1 (src/msm/mod.zig:24-30).
 *    Zig's AffinePoint is not an extern struct, so the shim marshals once per SRS (SURVEY H-f).
 *  - A handle must not be used from two host threads at once.  Host buffers are borrowed for the
 *    duration of the call only.
 */
#ifndef ZOLT_GPU_H
#define ZOLT_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZOLT_OK 0
#define ZOLT_ERR_CUDA 1  /* a CUDA runtime call or kernel failed (message on stderr) */
#define ZOLT_ERR_ARG 2   /* invalid argument (null pointer, size not a power of two, ...) */
#define ZOLT_ERR_STATE 3 /* call sequence violated (e.g. bind after the last round) */
#define ZOLT_ERR_NOGPU 4 /* no CUDA device: there is no CPU path */
#define ZOLT_ERR_PEER 5  /* a peer GPU never delivered its part of an exchange (timeout) */

typedef struct zolt_ctx zolt_ctx;     /* one device + one stream */
typedef struct zolt_bases zolt_bases; /* device-resident G1 bases (an SRS) */
typedef struct zolt_sc zolt_sc;       /* one sumcheck instance: k device-resident tables */
typedef struct zolt_peer_box zolt_peer_box; /* this rank's NVLink mailbox + the mapped mailboxes of its peers */

/* ---- context ------------------------------------------------------------------------------ */
/* `stream` is a cudaStream_t to launch on (borrowed), or NULL to create a private one. */
int zolt_gpu_init(int device, void* stream, zolt_ctx** out);
int zolt_gpu_shutdown(zolt_ctx* ctx);
int zolt_gpu_sync(zolt_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int zolt_gpu_launch_count(zolt_ctx* ctx, unsigned long long* out);
const char* zolt_gpu_version(void);
/* Tuning knobs (tests use them to force a code path; defaults are what bench.py measures).
 *   ZOLT_TUNE_SC_SMALL_PAIRS  rounds whose current table has at most this many pairs run on the
 *                             latency-oriented small-round kernel (default 2^14; 0 = always stream)
 *   ZOLT_TUNE_SC_BIND_F64     bit k set: the fused adjacent-pair passes bind table k on the FP64 pipe (exact 16 x 32-bit
 *                             limb products in doubles, csrc/field.cuh fp_bind2_fixed_f64) instead of the integer
 *                             multiplier.  Same values; default 0 (measured on B200: the passes are bound by instruction
 *                             issue, the idle FP64 pipe buys nothing -- kept as a parity-tested alternative)
 *   ZOLT_TUNE_PEER_TIMEOUT_MS how long a kernel waits for its peers' part of an exchange before the call fails with
 *                             ZOLT_ERR_PEER (default 4000).  Measured from the waiting kernel's own start, so ranks should be
 *                             aligned (zolt_peer_barrier) after anything slow and rank-local: a first multi-GiB workspace
 *                             allocation, zolt_msm_bases_precompute, a debugger */
#define ZOLT_TUNE_SC_SMALL_PAIRS 1
#define ZOLT_TUNE_SC_BIND_F64 2
#define ZOLT_TUNE_PEER_TIMEOUT_MS 3
/*   ZOLT_TUNE_SC_TAIL_PAIRS   rounds whose table has at most this many pairs run inside ONE resident kernel (a cluster of 8
 *                             CTAs) that takes each challenge from mapped host memory and writes each round polynomial back,
 *                             instead of one launch per round (default 2^12; 0 = one launch per round).  zolt_sc_bind posts the
 *                             challenge, zolt_sc_round_evals reads the answer: the call pattern does not change.
 *   ZOLT_TUNE_SC_TAIL_TIMEOUT_MS  how long that kernel waits for the next challenge before it retires by itself (default
 *                             10000); a challenge that arrives later is served by a regular launch, transparently */
/*   ZOLT_TUNE_MSM_SORT       0 (default): MSMs over a window-level table sort their bucket entries with the two-pass bin sort
 *                             from 2^18 points up; 1: always the atomic counting-sort scatter (the A/B switch of tools/time_msm_sort.py) */
#define ZOLT_TUNE_MSM_SORT 9
#define ZOLT_TUNE_SC_TAIL_PAIRS 4
#define ZOLT_TUNE_SC_TAIL_TIMEOUT_MS 5
int zolt_gpu_set_tuning(zolt_ctx* ctx, int key, long long value);

/* raw device buffers, so a host without CUDA bindings can keep tables resident */
int zolt_dev_alloc(zolt_ctx* ctx, size_t bytes, void** dptr);
int zolt_dev_free(zolt_ctx* ctx, void* dptr);
/* The same from the context's buffer cache (what the library's own temporaries use): a released buffer serves the next
 * request of a similar size without cudaMalloc / cudaFree -- milliseconds each for GiB-sized tables, and cudaFree
 * synchronises the device.  Only for buffers whose work is issued on the context's stream; `bytes` on release must be the
 * size asked for.  The cache is emptied by zolt_gpu_shutdown. */
int zolt_dev_alloc_cached(zolt_ctx* ctx, size_t bytes, void** dptr);
int zolt_dev_release_cached(zolt_ctx* ctx, void* dptr, size_t bytes);
int zolt_dev_upload(zolt_ctx* ctx, void* dptr, const void* hptr, size_t bytes);
int zolt_dev_download(zolt_ctx* ctx, void* hptr, const void* dptr, size_t bytes);
/* device-to-device copy, asynchronous on the context's stream */
int zolt_dev_copy(zolt_ctx* ctx, void* d_dst, const void* d_src, size_t bytes);

/* ---- field vectors ---------------------------------------------------------------------------
 * Element-wise ops on host vectors; replaces BatchOps.batchAdd/batchMul (src/field/mod.zig:1164-1280)
 * and doubles as the parity hook for F1-F7 (montgomeryMul :735-779, add :782-798, sub :801-816,
 * neg :944-947, inverse :955-983, fromMontgomery :642-645, toMontgomery :649-652). */
enum { ZOLT_FIELD_FR = 0, ZOLT_FIELD_FQ = 1 };
enum {
  ZOLT_OP_MUL = 0,
  ZOLT_OP_ADD = 1,
  ZOLT_OP_SUB = 2,
  ZOLT_OP_INV = 3,       /* b ignored; inverse(0) = 0 (batchInverse convention, :1243-1271) */
  ZOLT_OP_FROM_MONT = 4, /* b ignored */
  ZOLT_OP_TO_MONT = 5,   /* b ignored; a is any 256-bit integer (fromBytes semantics :625-639) */
  ZOLT_OP_NEG = 6        /* b ignored */
};
int zolt_field_batch_op(zolt_ctx* ctx, int field, int op, const uint64_t* a, const uint64_t* b,
                        uint64_t* out, size_t n);

/* ---- MSM -------------------------------------------------------------------------------------
 * zolt_msm == msm.MSM(F, Fp).compute(bases[offset..offset+n], scalars)  (src/msm/mod.zig:355-372);
 * callers: HyperKZG.commit (src/poly/commitment/mod.zig:249), Dory row commitments
 * (src/poly/commitment/dory.zig:663, :1024).  Bases are uploaded once per SRS and stay resident
 * (SetupParams.powers_of_tau_g1 is long-lived, commitment/mod.zig:122-140). */
int zolt_msm_bases_upload(zolt_ctx* ctx, const uint64_t* xy /* n*8 */, const uint8_t* inf /* n or NULL */,
                          size_t n, zolt_bases** out);
int zolt_msm_bases_free(zolt_bases* bases);
/* Optional, once per SRS: build the window-level table 2^(c*j) * base[i] (j < ceil(255/c)) on the device.
 * Every later MSM over these bases (any offset / length, zolt_msm*, zolt_msm_batch) then runs as ONE
 * window over the table: c = 20 -> 13 mixed additions per point instead of 16, no per-window bucket
 * reduction.  window_bits = 0 picks c from the length (20 from 2^20 points, 16 from 2^15, none below).
 * Costs (W-1)*c doublings per base and W*n*64 bytes of HBM (13.9 GB at 2^24); results are unchanged
 * (the same group element, canonical affine).  The SRS is long-lived in the reference
 * (SetupParams.powers_of_tau_g1, src/poly/commitment/mod.zig:122-140), which is what makes this pay. */
int zolt_msm_bases_precompute(zolt_bases* bases, int window_bits);
int zolt_msm_bases_len(const zolt_bases* bases, size_t* n);
int zolt_msm_bases_download(const zolt_bases* bases, size_t offset, size_t n, uint64_t* xy, uint8_t* inf);

int zolt_msm(zolt_ctx* ctx, const zolt_bases* bases, size_t offset, const uint64_t* scalars /* n*4, host */,
             size_t n, uint64_t out_xy[8], uint8_t* out_inf);
/* same, scalars already on the device (e.g. a bound sumcheck table or a quotient in HyperKZG.open) */
int zolt_msm_dev(zolt_ctx* ctx, const zolt_bases* bases, size_t offset, const void* d_scalars, size_t n,
                 uint64_t out_xy[8], uint8_t* out_inf);
/* k MSMs over the same n bases: BatchMSM.compute (src/msm/mod.zig:551-563), the shape of
 * DoryCommitmentScheme.computeRowCommitments (dory.zig:646-670) and HyperKZG.batchCommit
 * (commitment/mod.zig:558-570). scalars = k*n*4, outputs k*8 / k. */
int zolt_msm_batch(zolt_ctx* ctx, const zolt_bases* bases, const uint64_t* scalars, size_t k, size_t n,
                   uint64_t* out_xy, uint8_t* out_inf);
/* Multi-GPU: the un-normalised partial sum of this rank's point range as 16 u64 (extended Jacobian
 * X, Y, ZZ, ZZZ in Fq Montgomery form; ZZ == 0 encodes the identity), left in device memory at d_out
 * (128 bytes) so it can be all-gathered over NCCL; zolt_g1_sum_partials adds `count` such partials on
 * the device and normalises.  Mirrors ParallelMSM.compute's chunk/combine split
 * (src/msm/mod.zig:588-653). */
int zolt_msm_partial_dev(zolt_ctx* ctx, const zolt_bases* bases, size_t offset, const void* d_scalars,
                         size_t n, void* d_out /* 16 u64 */);
int zolt_g1_sum_partials(zolt_ctx* ctx, const void* d_partials /* count*16 u64 */, size_t count,
                         uint64_t out_xy[8], uint8_t* out_inf);

/* bases[i] = scalars[i] * G, G = (1,2): fixed-base generator used for SRS / workload generation.
 * With scalars = tau^i this is HyperKZG.setup's powers_of_tau_g1 (commitment/mod.zig:174-213). */
int zolt_g1_fixed_base_mul(zolt_ctx* ctx, const uint64_t* scalars /* n*4 Montgomery Fr */, size_t n,
                           zolt_bases** out);
/* HyperKZG.setup(max_degree): mock SRS with tau = 0x12345678 (commitment/mod.zig:189) built on device. */
int zolt_hyperkzg_setup(zolt_ctx* ctx, size_t max_degree, zolt_bases** out);
/* HyperKZG.open (commitment/mod.zig:261-324): fold the table variable by variable (half-split order); per level
 * the quotient hi - lo is committed (an MSM over the first len/2 powers) and the table is folded with point[i].
 * Everything stays on the device between levels.  Outputs: quotient commitments (affine, n_quotients of them:
 * num_vars unless the table has fewer than 2^num_vars entries) and final_eval (the fully folded value; for
 * num_vars == 0 the caller keeps `value`, as the reference does).  The _dev variant borrows a device table
 * (not modified). */
int zolt_hyperkzg_open(zolt_ctx* ctx, const zolt_bases* params, const uint64_t* evals /* n*4, host */, size_t n,
                       const uint64_t* point /* num_vars*4 */, int num_vars, uint64_t* quotients_xy /* num_vars*8 */,
                       uint8_t* quotients_inf /* num_vars */, int* n_quotients, uint64_t final_eval[4]);
int zolt_hyperkzg_open_dev(zolt_ctx* ctx, const zolt_bases* params, const void* d_evals, size_t n, const uint64_t* point,
                           int num_vars, uint64_t* quotients_xy, uint8_t* quotients_inf, int* n_quotients,
                           uint64_t final_eval[4]);

/* ---- Dory, G1 side, and the wire formats that feed the MSM (pairings / G2 / GT stay on the CPU) ---------------------------
 * zolt_msm_bases_upload_wire: n G1 points of 64 bytes each straight from a file into resident bases; the conversion to
 * Montgomery form, the identity detection and the curve check run on the device.
 *   ZOLT_G1_WIRE_BE      x || y big-endian standard integers: the raw-binary SRS (srs.zig:65-99, :256-306); all-zero =
 *                        identity; a point off y^2 = x^3 + 3 fails the call (ZOLT_ERR_ARG, *bad_index = first offender:
 *                        SRSError.PointNotOnCurve)
 *   ZOLT_G1_WIRE_LE      little-endian: a snarkjs PTAU tauG1 section as the reference reads it (srs.zig:616-660, :733-900)
 *   ZOLT_G1_WIRE_ARK_LE  little-endian, arkworks flag bits in the top of y cleared, no curve check: the Jolt Dory SRS
 *                        file (dory.zig:740-751, :828-880)
 * The container parsers (headers, section tables) are host glue: host/zolt_host.hpp srs::. */
#define ZOLT_G1_WIRE_BE 0
#define ZOLT_G1_WIRE_LE 1
#define ZOLT_G1_WIRE_ARK_LE 2
int zolt_msm_bases_upload_wire(zolt_ctx* ctx, const uint8_t* bytes /* n*64 */, size_t n, int format, zolt_bases** out,
                               size_t* bad_index /* may be NULL */);
/* A bases handle over device-resident points the caller owns (not copied, not freed): lets every MSM entry point run on a
 * vector that was produced on the device, e.g. Dory's folded row commitments v1 (dory.zig:1298-1299 E1+ / E1-). */
int zolt_msm_bases_wrap_dev(zolt_ctx* ctx, const void* d_xy /* n*64 */, const uint8_t* d_inf /* n or NULL */, size_t n,
                            zolt_bases** out);
int zolt_msm_bases_dev_ptrs(const zolt_bases* bases, const void** d_xy, const uint8_t** d_inf);
/* compressG1 (dory.zig:51-76, yIsPositive :162-175): 32 bytes per point, x little-endian standard form, flags 0x40 infinity /
 * 0x80 "y negative" in the top bits of the last byte. */
int zolt_g1_compress(zolt_ctx* ctx, const zolt_bases* pts, size_t offset, size_t n, uint8_t* out32 /* n*32 */);
/* v[i] += scalar * g[i] for i < n: the first-challenge update of a Dory reduce round (dory.zig:1277-1281) on resident vectors */
int zolt_g1_scale_add(zolt_ctx* ctx, zolt_bases* v, const zolt_bases* g, size_t n, const uint64_t scalar[4]);
/* The G1 half of DoryCommitmentScheme.commit (dory.zig:989-1042, computeRowCommitments :646-670): sigma = (num_vars + 1) / 2,
 * nu = num_vars - sigma with num_vars = floor(log2 len) (1 for a single entry); row r is the MSM of evals[r 2^sigma ..] over
 * g1_vec[0 .. 2^sigma): 2^nu MSMs over shared bases, run as ONE batched MSM.  Entries beyond 2^num_vars are ignored, as in
 * the reference.  rows_xy / rows_inf hold up to 2^nu results; *num_rows, *num_cols report the split. */
int zolt_dory_row_commitments(zolt_ctx* ctx, const zolt_bases* g1, const uint64_t* evals /* len*4, host */, size_t len,
                              uint64_t* rows_xy, uint8_t* rows_inf, size_t* num_rows, size_t* num_cols);
int zolt_dory_row_commitments_dev(zolt_ctx* ctx, const zolt_bases* g1, const void* d_evals, size_t len, uint64_t* rows_xy,
                                  uint8_t* rows_inf, size_t* num_rows, size_t* num_cols);

/* HyperKZG.batchOpen (commitment/mod.zig:607-732): evaluations[i] = evaluateMultilinear(polys[i], point) exactly as the
 * reference writes it (:788-817: index bit j <-> point[j]; the direct sum for at most 1024 entries and 10 variables, and
 * polys[i][0] for anything larger -- that fallback is reproduced, not corrected, so proofs stay byte-identical); the
 * batching challenge gamma = 0x9a8b7c6d * prod_i (point[i] + 11) (1 if zero, :631-638); the combined table
 * sum_i gamma^i polys[i] over the first table's length, built on the device; then the fold-and-commit loop of open().
 * With num_vars == 0 final_eval = sum_i gamma^i evaluations[i] and there are no quotients.  k == 0: final_eval 0, gamma 1.
 * Outputs: evaluations k*4, quotient commitments as in zolt_hyperkzg_open, final_eval, batching_challenge. */
int zolt_hyperkzg_batch_open(zolt_ctx* ctx, const zolt_bases* params, const uint64_t* const* polys /* k host tables */,
                             const size_t* lens /* k */, size_t k, const uint64_t* point, int num_vars, uint64_t* evaluations,
                             uint64_t* quotients_xy, uint8_t* quotients_inf, int* n_quotients, uint64_t final_eval[4],
                             uint64_t batching_challenge[4]);
int zolt_hyperkzg_batch_open_dev(zolt_ctx* ctx, const zolt_bases* params, const void* const* d_polys /* k device tables */,
                                 const size_t* lens, size_t k, const uint64_t* point, int num_vars, uint64_t* evaluations,
                                 uint64_t* quotients_xy, uint8_t* quotients_inf, int* n_quotients, uint64_t final_eval[4],
                                 uint64_t batching_challenge[4]);

/* ---- dense multilinear tables ------------------------------------------------------------------
 * Device-resident tables of 2^v Fr elements (32 B each, same layout as []F). */
/* DensePolynomial.bindLow (src/poly/mod.zig:160-175): dst[i] = src[2i] + r*(src[2i+1]-src[2i]).
 * dst may equal src (in place, like the reference) -- n is the CURRENT length, n/2 are written. */
int zolt_poly_bind_low(zolt_ctx* ctx, void* d_dst, const void* d_src, size_t n, const uint64_t r[4]);
/* DensePolynomial.bindFirst (src/poly/mod.zig:128-149): dst[i] = (1-r)*src[i] + r*src[i+n/2]. */
int zolt_poly_bind_first(zolt_ctx* ctx, void* d_dst, const void* d_src, size_t n, const uint64_t r[4]);
/* EqPolynomial.evals / evalsSliceWithScaling (src/poly/mod.zig:240-290): big-endian table
 * eq(r, x), r[0] <-> MSB of the index; scale may be NULL (= one).  d_out holds 2^n_vars elements. */
int zolt_eq_evals(zolt_ctx* ctx, const uint64_t* r /* n_vars*4 */, int n_vars, const uint64_t* scale,
                  void* d_out);
/* EqPlusOnePolynomial.mle(r, j) for every j in {0,1}^n_vars (src/poly/mod.zig:387-421; the table the Stage 3 provers build with
 * one mle per entry, src/zkvm/spartan/stage3_prover.zig:1878-1895): eq+1(r, j) = 1 iff j = r + 1 on the hypercube; r[0] <-> MSB
 * of the index like zolt_eq_evals.  d_out holds 2^n_vars elements. */
int zolt_eq_plus_one_evals(zolt_ctx* ctx, const uint64_t* r /* n_vars*4 */, int n_vars, void* d_out);
/* sum of all entries (initial sumcheck claim, src/subprotocols/mod.zig:306-309) */
int zolt_poly_sum(zolt_ctx* ctx, const void* d_src, size_t n, uint64_t out[4]);
/* BatchOps.innerProduct (src/field/mod.zig:1164-1280) over device tables: sum_i a[i] * b[i] */
int zolt_fr_inner_product_dev(zolt_ctx* ctx, const void* d_a, const void* d_b, size_t n, uint64_t out[4]);
/* Element-wise construction of device-resident tables (asynchronous on the context's stream):
 *   ZOLT_VEC_MUL            out = a o b
 *   ZOLT_VEC_EQ_AB_MINUS_C  out = e * (a * b - c)   JoltSpartanInterface's combined table (jolt_r1cs.zig:352-360)
 *   ZOLT_VEC_AXPY           out = out + s * a       gamma-power combination of batchOpen (commitment/mod.zig:648-656)
 *   ZOLT_VEC_AFFINE_INDEX   out[i] = s[0] + s[1] * i   UnmapPolynomial as a table (ram/raf_checking.zig:206-209): with it
 *                           RafEvaluationProver (:335-409) is the PROD2 combiner over (ra, unmap) with zolt_sc_set_claim
 * Unused operands may be NULL; s holds one scalar (AXPY) or two (AFFINE_INDEX). */
#define ZOLT_VEC_MUL 0
#define ZOLT_VEC_EQ_AB_MINUS_C 1
#define ZOLT_VEC_AXPY 2
#define ZOLT_VEC_AFFINE_INDEX 3
/*   ZOLT_VEC_LIN2           out = s[0] * a + s[1] * b   ra = gamma rs1_ra + gamma^2 rs2_ra (stage4_gruen_prover.zig:211, :224)
 *   ZOLT_VEC_STAGE4_TERM    out = e * (a * b + c * (b + out))   one term of computeInputClaim (:385-430): e = eq, a = ra,
 *                           b = val, c = wa, out holds inc on entry */
#define ZOLT_VEC_LIN2 4
#define ZOLT_VEC_STAGE4_TERM 5
int zolt_fr_vec_dev(zolt_ctx* ctx, int op, void* d_out, const void* d_a, const void* d_b, const void* d_c, const void* d_e,
                    const uint64_t* s /* 4, AXPY only */, size_t n);

/* d_out[x_lo] = scale * sum_{x_hi < n_hi} d_in[x_lo + x_hi * n_lo] * d_w[x_hi] for x_lo < n_lo (scale NULL = one): the Q buffers of
 * the prefix-suffix provers, a trace column folded over the high half of the cycle index against the suffix table
 * (ShiftPrefixSuffixProver.init, src/zkvm/spartan/stage3_prover.zig:1053-1112; RegistersPrefixSuffixProver.init, :2262-2300). */
int zolt_fr_fold_high_dev(zolt_ctx* ctx, void* d_out, const void* d_in, const void* d_w, size_t n_lo, size_t n_hi,
                          const uint64_t* scale /* 4 or NULL */);

/* LassoProver address rounds over device-resident eq_evals (n field elements, updated in place) and the 128-bit lookup indices
 * (n x 16 bytes, little-endian), src/zkvm/lasso/prover.zig: one pass applies the previous challenge -- eq_evals[j] *= r if bit
 * bind_bit of index j is set, else (1 - r) (receiveChallenge :380-400; bind_bit < 0: nothing to apply) -- and returns the partition
 * sums of computeAddressRoundPoly (:265-310) for bit sum_bit: sums_out = [sum_0, sum_1] (sum_bit < 0: [sum of the table, 0]).
 * sum_0 + sum_1 is the reference's claim after the bind (:402-407).  The cycle rounds (:312-364, :421-452) are the
 * ZOLT_SC_SUM / ZOLT_PAIR_HALF instance over the same table. */
int zolt_lasso_addr_round(zolt_ctx* ctx, void* d_eq_evals, const void* d_indices, size_t n, int bind_bit, const uint64_t* r /* 4 */,
                          int sum_bit, uint64_t sums_out[8]);

/* dst[i] = src[((i / d1) % m1) * s1 + ((i / d2) % m2) * s2] for i < n: re-layouts between the phases of a prover --
 * transpose [k][j] -> [j][k] (d1 = 1, m1 = K, s1 = T, d2 = K, m2 = T, s2 = 1), a vector replicated over rows (d1 = 1,
 * m1 = len, s1 = 1, s2 = 0), every element repeated (d1 = K, m1 = len, s1 = 1, s2 = 0).  m = 0 disables a term. */
int zolt_fr_index_map_dev(zolt_ctx* ctx, void* d_dst, const void* d_src, size_t n, size_t d1, size_t m1, size_t s1, size_t d2,
                          size_t m2, size_t s2);

/* ---- sumcheck prover inner loop ----------------------------------------------------------------
 * One instance = k dense tables bound in lock-step.  The call pattern is the reference's:
 *   round_evals -> (host transcript) -> bind -> round_evals -> ...
 * (SumcheckInstance.computeRoundPolyFn / bindChallengeFn, src/zkvm/batched_sumcheck.zig:34-74;
 *  prover.SumcheckInstance vtable, src/zkvm/prover.zig:30-69).
 * zolt_sc_bind(r) launches ONE fused kernel that binds every table to r and accumulates the next
 * round's evaluations in the same pass over HBM; the following zolt_sc_round_evals only reads the
 * 128-byte result. */
enum {
  ZOLT_PAIR_ADJACENT = 0, /* LowToHigh: pairs (2i, 2i+1)   -- bindLow order */
  ZOLT_PAIR_HALF = 1      /* HighToLow: pairs (i, i+n/2)   -- bindFirst order */
};
enum {
  /* k=1. evals = [sum lo, sum hi].  Sumcheck.Prover.nextRound (src/subprotocols/mod.zig:69-109),
   * JoltSpartanInterface.computeRoundPolynomial (src/zkvm/r1cs/jolt_r1cs.zig:413-454), stages 5/6
   * (src/zkvm/prover.zig:905-946). */
  ZOLT_SC_SUM = 0,
  /* k=2. evals at X=0,1,2,3 of sum a(X)*b(X).  ValFinalProver (src/zkvm/ram/val_final.zig:150-184);
   * the Eq*MLE shape of BASELINE config 2. */
  ZOLT_SC_PROD2 = 1,
  /* k=3. evals at X=0..3 of sum a*b*c.  ValEvaluationProver (src/zkvm/ram/val_evaluation.zig:554-604). */
  ZOLT_SC_PROD3 = 2,
  /* k=4 (eq, a, b, c). sum eq*a*(b-c).  OutputSumcheckProver (src/zkvm/ram/output_check.zig:375-433). */
  ZOLT_SC_EQ_A_BMC = 3,
  /* k=4 (eq, az, bz, cz). sum eq*(az*bz - cz): the Spartan-style shape of BASELINE config 3
   * (elementwise form at src/zkvm/r1cs/jolt_r1cs.zig:352-360). */
  ZOLT_SC_EQ_AB_MINUS_C = 4,
  /* k=2 (left, right) + split-eq weights.  evals = [t0, t_inf]:
   *   t0 = sum_g E_out[x_out] E_in[x_in] l_lo r_lo, t_inf = sum_g E_out E_in (l_hi-l_lo)(r_hi-r_lo)
   * ProductVirtualRemainderProver (src/zkvm/spartan/product_remainder.zig:269-322),
   * StreamingOuterProver.buildTPrimePoly (src/zkvm/spartan/streaming_outer.zig:411-457).
   * The host finishes with GruenSplitEqPolynomial.computeCubicRoundPoly (src/poly/split_eq.zig:353). */
  ZOLT_SC_GRUEN_AB = 5,
  /* k=3 (az, bz, cz) + split-eq weights: t0 = sum w (az0 bz0 - cz0), t_inf = sum w daz dbz. */
  ZOLT_SC_GRUEN_AB_MINUS_C = 6,
  /* Stage4GruenProver (registers read/write checking, src/zkvm/spartan/stage4_gruen_prover.zig), body
   * C(X) = ra(X) val(X) + wa(X) (val(X) + inc(X)) over K x T tables laid out [k][j] with the row stride equal to the
   * CURRENT cycle length, so that the reference's per-row pairs (k*T + 2i, k*T + 2i + 1) are the adjacent pairs of the flat
   * table.  Both return raw sums (no claim handling on the device; the O(1) completion is host glue).
   * k=4 (ra, wa, val, inc replicated over k) + split-eq weights over the CYCLE bits only (zolt_sc_set_split_eq with
   * n_tau = log_n - 7: the pair's row k is masked out of the weight index) -> [q(0), q_X2]
   * = sum_i E_out E_in sum_k [C(0), leading coefficient of C]  (phase1ComputeMessage, :561-757). */
  ZOLT_SC_STAGE4_GRUEN = 7,
  /* k=5 (eq, ra, wa, val, inc), all of the same length -> [p(0), p(2), p(3)] of sum eq(X) C(X): phases 2 and 3
   * (:764-1015) on tables transposed to [j][k] with eq and inc repeated over k -- an address round binds them to
   * themselves (the round polynomial then has degree 2 and p(3) is not used), a cycle round (k exhausted) is the
   * dense-eq cubic. */
  ZOLT_SC_STAGE4_DENSE = 8,
  /* k=5 (w, a, b, c, d) -> [p(0), p(2), p(3)] of sum w(X) (a(X) b(X) + c(X) d(X)): one operand side of InstructionInputProver
   * (src/zkvm/spartan/stage3_prover.zig:2029-2105; right = is_rs2 * rs2 + is_imm * imm, left = is_rs1 * rs1 + is_pc * pc under
   * the weight eq_outer + gamma^2 eq_product -- the host adds right + gamma * left, exact in the field).  Adjacent pairs only. */
  ZOLT_SC_EQ_DOT2 = 9
};

/* tables: k host pointers (n*4 u64 each), n = 2^log_n.  n_extra further tables are bound in
 * lock-step but do not enter the round polynomial (val_init in output_check.zig:454-476). */
int zolt_sc_create(zolt_ctx* ctx, int combiner, int pairing, int log_n, int n_tables, int n_extra,
                   const uint64_t* const* tables, zolt_sc** out);
/* same, from device-resident tables: BORROWED read-only (never written, no copy); they must stay valid
 * until the first zolt_sc_bind has completed (or the instance is freed before that). */
int zolt_sc_create_dev(zolt_ctx* ctx, int combiner, int pairing, int log_n, int n_tables, int n_extra,
                       const void* const* d_tables, zolt_sc** out);
/* Gruen combiners only: tau (n_tau*4, n_tau == log_n; ZOLT_SC_STAGE4_GRUEN: n_tau <= log_n, the weights then depend on the
 * low n_tau - 1 bits of the pair index only) -> builds the E_out / E_in prefix tables of
 * GruenSplitEqPolynomial.initWithScaling on the device (src/poly/split_eq.zig:63-183; the same tables as
 * src/zkvm/spartan/gruen_eq.zig:296-334). */
int zolt_sc_set_split_eq(zolt_sc* sc, const uint64_t* tau, int n_tau);
/* number of evaluations zolt_sc_round_evals writes (2, 3 or 4) */
int zolt_sc_num_evals(const zolt_sc* sc, int* out);
int zolt_sc_round_evals(zolt_sc* sc, uint64_t* evals_out /* num_evals*4 */);
int zolt_sc_bind(zolt_sc* sc, const uint64_t challenge[4]);
/* current (bound) length of the tables and a copy of table t */
int zolt_sc_len(const zolt_sc* sc, size_t* n);
int zolt_sc_download(zolt_sc* sc, int table, uint64_t* dst /* len*4 */);
/* device pointer of table t's current buffer (valid until the next bind) */
int zolt_sc_table_dev(zolt_sc* sc, int table, const void** d_ptr);
/* Optional: the instance's input claim (SumcheckInstance.input_claim, src/zkvm/batched_sumcheck.zig:34-74).
 * Every round after the first derives p(1) from the sumcheck identity p(0) + p(1) = claim (claim =
 * previous round polynomial at the challenge, tracked on the device); with the input claim set the first
 * round does too.  Must be called before the first zolt_sc_round_evals.  The evaluations are the same
 * canonical field elements either way. */
int zolt_sc_set_claim(zolt_sc* sc, const uint64_t claim[4]);
/* Multi-GPU (hypercube shards, one instance per rank).  After zolt_sc_set_sharded(sc, 1) a round is:
 *   zolt_sc_round_partials_dev -> this rank's 4 raw partial sums, left on the device (16 u64) for the
 *                                 all-gather;
 *   zolt_sc_finish_partials    -> folds the `rows` gathered partials mod r on the device, completes the
 *                                 round polynomial (p(1) from the claim, p(3) from the third difference)
 *                                 and returns the evaluations; every rank calls it with the same rows;
 *   zolt_sc_bind               -> as usual. */
int zolt_sc_set_sharded(zolt_sc* sc, int on);
int zolt_sc_round_partials_dev(zolt_sc* sc, const void** d_ptr);
int zolt_sc_finish_partials(zolt_sc* sc, const void* d_rows /* rows*16 u64 */, size_t rows,
                            uint64_t* evals_out /* num_evals*4 */);
int zolt_fr_sum_rows(zolt_ctx* ctx, const void* d_rows, size_t rows, size_t cols, uint64_t* out /* cols*4 */);
/* Multi-GPU, fused exchange (preferred): after zolt_sc_set_peers(sc, box) every round kernel pushes its 4 raw
 * sums into the mailbox of every rank over NVLink and folds all ranks' rows itself, so zolt_sc_round_evals /
 * zolt_sc_bind behave exactly as on one GPU while each rank holds only its hypercube block.  Every rank
 * must make the same sequence of calls.  zolt_sc_gather_tail: once the local block is down to ZOLT_SC_TAIL_ENTRIES
 * entries per table or fewer, all-gather the blocks over NVLink: d_tables_out[t] then points at the GLOBAL table t
 * (*n_out = world * local length, blocked order) inside this rank's mailbox, valid until the box's next but one
 * gather; every rank runs the remaining rounds on those tables redundantly, without further exchanges (small rounds
 * are latency-bound: one exchange less per round matters more than the duplicated arithmetic). */
#define ZOLT_SC_TAIL_ENTRIES 1024
int zolt_sc_set_peers(zolt_sc* sc, zolt_peer_box* box);
int zolt_sc_gather_tail(zolt_sc* sc, const void** d_tables_out /* k+n_extra device pointers */, size_t* n_out);
int zolt_sc_free(zolt_sc* sc);

/* ---- NVLink peer mailbox (multi-GPU exchange without a collective library) ---------------------------
 * One process per GPU.  zolt_peer_box_create allocates this rank's mailbox and returns its 64-byte CUDA IPC
 * handle; the caller exchanges the handles of all ranks by any means (bench.py: torch.distributed
 * all_gather at start-up) and passes them, rank-ordered, to zolt_peer_box_connect, which maps every peer's
 * mailbox (P2P over NVLink/NVSwitch).  From then on kernels exchange small results by storing straight into
 * the peers' mailboxes (csrc/peer.cuh); a peer that never shows up makes the waiting call fail with
 * ZOLT_ERR_PEER after 4 s.  world <= 8.
 *   zolt_peer_allgather    n_fr (<= 8) field-sized elements per rank; d_out is [rank][n_fr] (transpose 0) or
 *                          [n_fr][rank] (transpose 1), device memory; synchronises the stream.
 *   zolt_msm_sharded_dev   MSM over point-range shards (ParallelMSM.compute's split, src/msm/mod.zig:608-653,
 *                          across GPUs): this rank's partial over its bases/scalars, exchanged and summed on
 *                          the device by one kernel; every rank returns the same affine result. */
int zolt_peer_box_create(zolt_ctx* ctx, zolt_peer_box** out, unsigned char handle_out[64]);
int zolt_peer_box_connect(zolt_peer_box* box, const unsigned char* handles /* world*64 */, int rank, int world);
/* ranks that live in one process (one context each): address the mailboxes directly instead of through IPC handles */
int zolt_peer_box_connect_local(zolt_peer_box* box, zolt_peer_box* const* boxes /* world, rank-ordered */, int rank, int world);
int zolt_peer_box_free(zolt_peer_box* box);
int zolt_peer_allgather(zolt_peer_box* box, const void* d_src, int n_fr, void* d_out, int transpose);
/* stream-ordered barrier across the ranks (an empty exchange); asynchronous */
int zolt_peer_barrier(zolt_peer_box* box);
int zolt_msm_sharded_dev(zolt_ctx* ctx, const zolt_bases* bases, size_t offset, const void* d_scalars, size_t n,
                         zolt_peer_box* box, uint64_t out_xy[8], uint8_t* out_inf);
/* the same from HOST scalars (this rank's range); the upload overlaps the multiplication as in zolt_msm */
int zolt_msm_sharded(zolt_ctx* ctx, const zolt_bases* bases, size_t offset, const uint64_t* scalars /* n*4, host */, size_t n,
                     zolt_peer_box* box, uint64_t out_xy[8], uint8_t* out_inf);

/* ---- one process, several devices --------------------------------------------------------------------
 * `zolt prove` is ONE process (src/zkvm/mod.zig:366-462).  zolt_gpu_init_multi (SURVEY 8b's proposal) creates a context and
 * a mailbox per entry of device_ids, connects the mailboxes directly (peer access between distinct devices, no IPC) and
 * hands out the per-rank handles, which are ordinary zolt_ctx / zolt_peer_box: one host thread per device then drives
 * zolt_msm_sharded*, zolt_sc_set_peers or the C++ round loop exactly as a rank of the one-process-per-GPU deployment does.
 * A device may be listed more than once (ranks share it): that is how a one-GPU box exercises this path.
 * zolt_multi_msm: MSM.compute over all bases = ParallelMSM.compute's point-range split (src/msm/mod.zig:588-653) with one GPU
 * and one host thread per chunk; bases[r] is rank r's contiguous range (resident on its device), scalars the whole vector. */
typedef struct zolt_multi zolt_multi;
int zolt_gpu_init_multi(const int* device_ids, int n_devices, zolt_multi** out);
int zolt_multi_shutdown(zolt_multi* m);
int zolt_multi_size(const zolt_multi* m, int* n_devices);
int zolt_multi_rank(zolt_multi* m, int rank, zolt_ctx** ctx, zolt_peer_box** box);
int zolt_multi_msm(zolt_multi* m, zolt_bases* const* bases /* n_devices */, const uint64_t* scalars /* n_total*4, host */,
                   size_t n_total, uint64_t out_xy[8], uint8_t* out_inf);
/* zolt_msm_partial_dev from HOST scalars (upload overlapped with the multiplication as in zolt_msm) */
int zolt_msm_partial(zolt_ctx* ctx, const zolt_bases* bases, size_t offset, const uint64_t* scalars, size_t n, void* d_out /* 16 u64 */);

/* ---- diagnostics ---------------------------------------------------------------------------------
 * Integer-pipe micro-benchmarks (the MSM roofline denominator).  `kind`: 0 = IMAD (3-reg),
 * 1 = IMAD.WIDE.U32, 2 = Fr Montgomery products, 3 = Fq Montgomery products, 4 = IMAD immediate form,
 * 5 = FP64 FMA, 6 = FP64 FMA interleaved 1:1 with IMAD (ops/s counts the FP64 FMAs only).
 * Returns ops per second (for kind 2/3: field products per second). */
int zolt_bench_int_pipe(zolt_ctx* ctx, int kind, int iters, double* ops_per_s);
/* same with an explicit launch shape (threads per block, resident blocks per SM): occupancy sweeps */
int zolt_bench_int_pipe_cfg(zolt_ctx* ctx, int kind, int iters, int threads, int blocks_per_sm,
                            double* ops_per_s);
/* Per-kernel device timings (CUDA events on the context's stream, the stream the kernels are launched
 * on).  zolt_prof_enable(ctx, 1) clears the log and starts recording one (tag, ms) entry per phase of
 * every subsequent MSM / sumcheck call (up to 96 entries); zolt_prof_read synchronises and copies them.
 * Tags: 1 digits, 2 scan, 3 scatter, 4 bucket accumulation (level 0), 5 edge levels, 6 bucket
 * reduction, 7 normalisation; 16 sumcheck round-0 evaluation, 17 fused bind+evaluate round. */
int zolt_prof_enable(zolt_ctx* ctx, int on);
int zolt_prof_read(zolt_ctx* ctx, int* tags, float* ms, int max_entries, int* n_entries);
/* device timing helpers on the context's stream (CUDA events) */
int zolt_timer_start(zolt_ctx* ctx);
int zolt_timer_stop_ms(zolt_ctx* ctx, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* ZOLT_GPU_H */
