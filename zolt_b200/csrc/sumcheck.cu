// Sumcheck prover inner loop: round-polynomial evaluation over k dense Fr tables, fused with the
// bind of the previous round's challenge so that every round is one pass over HBM.
//
// Replaces the O(N) loops of (reference, /root/reference):
//   src/subprotocols/mod.zig:69-122             Sumcheck.Prover.nextRound / receiveChallenge
//   src/zkvm/r1cs/jolt_r1cs.zig:413-489         JoltSpartanInterface round / bind
//   src/zkvm/ram/val_final.zig:150-206          ValFinalProver (dense degree 2)
//   src/zkvm/ram/val_evaluation.zig:554-640     ValEvaluationProver (dense degree 3)
//   src/zkvm/ram/output_check.zig:375-479       OutputSumcheckProver (eq * io * (vf - vio))
//   src/zkvm/spartan/product_remainder.zig:269-378  Gruen split-eq degree 3
//   src/poly/mod.zig:128-175, :240-290          bindFirst / bindLow / EqPolynomial.evals
//
// HBM layout: each table is an array of 32-byte Fr elements exactly as the reference's []F
// (Montgomery limbs), index bit j <-> variable j.  Every instance owns two buffers per table
// (sizes N and N/2) used as a ping-pong, so a fused bind+eval pass reads the current table once
// and writes the half-size table once: 128*k*N algorithmic bytes for a whole sumcheck.
#include <cstring>
#include <new>

#include "common.cuh"
#include "field.cuh"

using namespace zolt;

namespace {

constexpr int kMaxTables = 8;
constexpr int kScThreads = 256;

struct ScPtrs {
  const Fr* src[kMaxTables];
  Fr* dst[kMaxTables];
};

struct SplitEqDev {
  const Fr* e_out;  // level `out_bits` table of E_out_vec
  const Fr* e_in;   // level `in_bits` table of E_in_vec
  int in_bits;
};

__device__ __forceinline__ Fr ld_fr(const Fr* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  Fr r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}

__device__ __forceinline__ void st_fr(Fr* p, const Fr& x) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
  q[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}

__device__ __forceinline__ Fr fr_add(const Fr& a, const Fr& b) { return fp_add<FrParams>(a, b); }
__device__ __forceinline__ Fr fr_sub(const Fr& a, const Fr& b) { return fp_sub<FrParams>(a, b); }
__device__ __forceinline__ Fr fr_mul(const Fr& a, const Fr& b) { return fp_mul<FrParams>(a, b); }

// lo + r*(hi - lo): DensePolynomial.bindLow's form (src/poly/mod.zig:171); equal, as a canonical
// field element, to the (1-r)*lo + r*hi form used at val_evaluation.zig:626-631.
__device__ __forceinline__ Fr bind_pair(const Fr& lo, const Fr& hi, const Fr& r) {
  return fr_add(lo, fr_mul(r, fr_sub(hi, lo)));
}

// ---- combiners: accumulate one pair (lo[], hi[]) into the running evaluations -----------------------
template <int COMB> struct CombTraits;
template <> struct CombTraits<ZOLT_SC_SUM> { static constexpr int K = 1, NE = 2; static constexpr bool GRUEN = false; };
template <> struct CombTraits<ZOLT_SC_PROD2> { static constexpr int K = 2, NE = 4; static constexpr bool GRUEN = false; };
template <> struct CombTraits<ZOLT_SC_PROD3> { static constexpr int K = 3, NE = 4; static constexpr bool GRUEN = false; };
template <> struct CombTraits<ZOLT_SC_EQ_A_BMC> { static constexpr int K = 4, NE = 4; static constexpr bool GRUEN = false; };
template <> struct CombTraits<ZOLT_SC_EQ_AB_MINUS_C> { static constexpr int K = 4, NE = 4; static constexpr bool GRUEN = false; };
template <> struct CombTraits<ZOLT_SC_GRUEN_AB> { static constexpr int K = 2, NE = 2; static constexpr bool GRUEN = true; };
template <> struct CombTraits<ZOLT_SC_GRUEN_AB_MINUS_C> { static constexpr int K = 3, NE = 2; static constexpr bool GRUEN = true; };

template <int COMB>
__device__ __forceinline__ void accumulate_pair(Fr* acc, const Fr* lo, const Fr* hi, const Fr& w) {
  if constexpr (COMB == ZOLT_SC_SUM) {
    acc[0] = fr_add(acc[0], lo[0]);
    acc[1] = fr_add(acc[1], hi[0]);
  } else if constexpr (COMB == ZOLT_SC_PROD2) {
    // a(X) = lo + X*(hi-lo); evaluations at X = 0,1,2,3 by repeated addition of the slope
    Fr da = fr_sub(hi[0], lo[0]), db = fr_sub(hi[1], lo[1]);
    acc[0] = fr_add(acc[0], fr_mul(lo[0], lo[1]));
    acc[1] = fr_add(acc[1], fr_mul(hi[0], hi[1]));
    Fr a = fr_add(hi[0], da), b = fr_add(hi[1], db);
    acc[2] = fr_add(acc[2], fr_mul(a, b));
    a = fr_add(a, da);
    b = fr_add(b, db);
    acc[3] = fr_add(acc[3], fr_mul(a, b));
  } else if constexpr (COMB == ZOLT_SC_PROD3 || COMB == ZOLT_SC_EQ_A_BMC) {
    Fr a0 = lo[0], a1 = hi[0], b0 = lo[1], b1 = hi[1], c0, c1;
    if constexpr (COMB == ZOLT_SC_PROD3) {
      c0 = lo[2];
      c1 = hi[2];
    } else {  // eq * a * (b - c)
      c0 = fr_sub(lo[2], lo[3]);
      c1 = fr_sub(hi[2], hi[3]);
    }
    Fr da = fr_sub(a1, a0), db = fr_sub(b1, b0), dc = fr_sub(c1, c0);
    acc[0] = fr_add(acc[0], fr_mul(fr_mul(a0, b0), c0));
    acc[1] = fr_add(acc[1], fr_mul(fr_mul(a1, b1), c1));
    Fr a = fr_add(a1, da), b = fr_add(b1, db), c = fr_add(c1, dc);
    acc[2] = fr_add(acc[2], fr_mul(fr_mul(a, b), c));
    a = fr_add(a, da);
    b = fr_add(b, db);
    c = fr_add(c, dc);
    acc[3] = fr_add(acc[3], fr_mul(fr_mul(a, b), c));
  } else if constexpr (COMB == ZOLT_SC_EQ_AB_MINUS_C) {
    // eq(X) * (az(X)*bz(X) - cz(X))
    Fr e = lo[0], a = lo[1], b = lo[2], c = lo[3];
    Fr de = fr_sub(hi[0], e), da = fr_sub(hi[1], a), db = fr_sub(hi[2], b), dc = fr_sub(hi[3], c);
    acc[0] = fr_add(acc[0], fr_mul(e, fr_sub(fr_mul(a, b), c)));
#pragma unroll
    for (int x = 1; x < 4; x++) {
      e = fr_add(e, de);
      a = fr_add(a, da);
      b = fr_add(b, db);
      c = fr_add(c, dc);
      acc[x] = fr_add(acc[x], fr_mul(e, fr_sub(fr_mul(a, b), c)));
    }
  } else if constexpr (COMB == ZOLT_SC_GRUEN_AB) {
    Fr p0 = fr_mul(lo[0], lo[1]);
    Fr slope = fr_mul(fr_sub(hi[0], lo[0]), fr_sub(hi[1], lo[1]));
    acc[0] = fr_add(acc[0], fr_mul(w, p0));
    acc[1] = fr_add(acc[1], fr_mul(w, slope));
  } else {  // ZOLT_SC_GRUEN_AB_MINUS_C
    Fr p0 = fr_sub(fr_mul(lo[0], lo[1]), lo[2]);
    Fr slope = fr_mul(fr_sub(hi[0], lo[0]), fr_sub(hi[1], lo[1]));
    acc[0] = fr_add(acc[0], fr_mul(w, p0));
    acc[1] = fr_add(acc[1], fr_mul(w, slope));
  }
}

// ---- block / grid reduction of NE field accumulators -------------------------------------------------
__device__ __forceinline__ Fr shfl_down_fr(const Fr& x, int delta) {
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_down_sync(0xffffffffu, x.v[i], delta);
  return r;
}

// Sum over the block; result valid in thread 0.  smem: NE * 32 Fr (enough for 32 warps).
template <int NE>
__device__ __forceinline__ void block_reduce(Fr* acc, Fr* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int e = 0; e < NE; e++) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc[e] = fr_add(acc[e], shfl_down_fr(acc[e], d));
  }
  if (lane == 0) {
#pragma unroll
    for (int e = 0; e < NE; e++) smem[e * 32 + warp] = acc[e];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int e = 0; e < NE; e++) {
      Fr x = lane < nwarps ? smem[e * 32 + lane] : Fr::zero();
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) x = fr_add(x, shfl_down_fr(x, d));
      acc[e] = x;
    }
  }
  __syncthreads();
}

// Each block publishes its partial; the last block to arrive folds all partials and writes the
// round evaluations to device memory and to the mapped host mailbox.
template <int NE>
__device__ __forceinline__ void grid_finish(Fr* acc, Fr* smem, Fr* partials, unsigned* ticket, Fr* d_evals,
                                            Fr* h_evals) {
  __shared__ bool is_last;
  block_reduce<NE>(acc, smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int e = 0; e < NE; e++) st_fr(&partials[(size_t)blockIdx.x * NE + e], acc[e]);
    __threadfence();
    unsigned t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  Fr tot[NE];
#pragma unroll
  for (int e = 0; e < NE; e++) tot[e] = Fr::zero();
  for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
    for (int e = 0; e < NE; e++) {
      const uint4* q = reinterpret_cast<const uint4*>(&partials[(size_t)b * NE + e]);
      uint4 lo4 = __ldcg(q), hi4 = __ldcg(q + 1);  // bypass L1: written by other SMs
      Fr x;
      x.v[0] = lo4.x; x.v[1] = lo4.y; x.v[2] = lo4.z; x.v[3] = lo4.w;
      x.v[4] = hi4.x; x.v[5] = hi4.y; x.v[6] = hi4.z; x.v[7] = hi4.w;
      tot[e] = fr_add(tot[e], x);
    }
  }
  block_reduce<NE>(tot, smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int e = 0; e < NE; e++) {
      st_fr(&d_evals[e], tot[e]);
      st_fr(&h_evals[e], tot[e]);
    }
    *ticket = 0;
    __threadfence_system();
  }
}

// ---- the round kernel ------------------------------------------------------------------------------
// FUSED == false: evaluate the round polynomial of the current tables (src, length n).
// FUSED == true : bind src (length n) to r into dst (length n/2) and evaluate the NEXT round's
//                 polynomial over dst in the same pass.  n_extra tables after the K combiner tables
//                 are bound only.
template <int COMB, int PAIR, bool FUSED>
__global__ void __launch_bounds__(kScThreads)
sc_round_kernel(ScPtrs t, int n_extra, size_t n, Fr r, SplitEqDev sq, Fr* partials, unsigned* ticket,
                Fr* d_evals, Fr* h_evals) {
  constexpr int K = CombTraits<COMB>::K;
  constexpr int NE = CombTraits<COMB>::NE;
  __shared__ Fr smem[NE * 32];
  Fr acc[NE];
#pragma unroll
  for (int e = 0; e < NE; e++) acc[e] = Fr::zero();

  const size_t n_cur = FUSED ? n / 2 : n;  // length of the table the evaluations are taken over
  const size_t pairs = n_cur / 2;
  const size_t half_old = n / 2;           // HALF pairing: distance between bind partners in src
  const size_t half_cur = n_cur / 2;       // HALF pairing: distance between eval partners

  for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < pairs;
       g += (size_t)gridDim.x * blockDim.x) {
    const size_t i_lo = (PAIR == ZOLT_PAIR_ADJACENT) ? 2 * g : g;
    const size_t i_hi = (PAIR == ZOLT_PAIR_ADJACENT) ? 2 * g + 1 : g + half_cur;
    Fr lo[K], hi[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
      if constexpr (FUSED) {
        const Fr* s = t.src[k];
        Fr a0, a1, b0, b1;
        if constexpr (PAIR == ZOLT_PAIR_ADJACENT) {
          a0 = ld_fr(s + 2 * i_lo); a1 = ld_fr(s + 2 * i_lo + 1);
          b0 = ld_fr(s + 2 * i_hi); b1 = ld_fr(s + 2 * i_hi + 1);
        } else {
          a0 = ld_fr(s + i_lo); a1 = ld_fr(s + i_lo + half_old);
          b0 = ld_fr(s + i_hi); b1 = ld_fr(s + i_hi + half_old);
        }
        lo[k] = bind_pair(a0, a1, r);
        hi[k] = bind_pair(b0, b1, r);
        st_fr(t.dst[k] + i_lo, lo[k]);
        st_fr(t.dst[k] + i_hi, hi[k]);
      } else {
        lo[k] = ld_fr(t.src[k] + i_lo);
        hi[k] = ld_fr(t.src[k] + i_hi);
      }
    }
    if constexpr (FUSED) {
      for (int k = K; k < K + n_extra; k++) {
        const Fr* s = t.src[k];
        Fr a0, a1, b0, b1;
        if constexpr (PAIR == ZOLT_PAIR_ADJACENT) {
          a0 = ld_fr(s + 2 * i_lo); a1 = ld_fr(s + 2 * i_lo + 1);
          b0 = ld_fr(s + 2 * i_hi); b1 = ld_fr(s + 2 * i_hi + 1);
        } else {
          a0 = ld_fr(s + i_lo); a1 = ld_fr(s + i_lo + half_old);
          b0 = ld_fr(s + i_hi); b1 = ld_fr(s + i_hi + half_old);
        }
        st_fr(t.dst[k] + i_lo, bind_pair(a0, a1, r));
        st_fr(t.dst[k] + i_hi, bind_pair(b0, b1, r));
      }
    }
    Fr w;
    if constexpr (CombTraits<COMB>::GRUEN) {
      // g = (x_out << in_bits) | x_in  (src/zkvm/spartan/product_remainder.zig:293-296)
      const size_t x_out = g >> sq.in_bits, x_in = g & ((size_t(1) << sq.in_bits) - 1);
      w = fr_mul(ld_fr(sq.e_out + x_out), ld_fr(sq.e_in + x_in));
    }
    accumulate_pair<COMB>(acc, lo, hi, w);
  }
  grid_finish<NE>(acc, smem, partials, ticket, d_evals, h_evals);
}

// Final bind (n == 2 -> 1) and stand-alone binds: dst[i] = bind(src pair i).
template <int PAIR>
__global__ void __launch_bounds__(kScThreads) bind_kernel(Fr* dst, const Fr* src, size_t n, Fr r) {
  const size_t half = n / 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < half; i += (size_t)gridDim.x * blockDim.x) {
    Fr lo = (PAIR == ZOLT_PAIR_ADJACENT) ? ld_fr(src + 2 * i) : ld_fr(src + i);
    Fr hi = (PAIR == ZOLT_PAIR_ADJACENT) ? ld_fr(src + 2 * i + 1) : ld_fr(src + i + half);
    st_fr(dst + i, bind_pair(lo, hi, r));
  }
}

// In-place adjacent-pair bind needs a staging pass when dst == src; bindLow's in-place contract is
// honoured by the host wrapper through a scratch buffer.

// Sum of a table (initial claim).
__global__ void __launch_bounds__(kScThreads) sum_kernel(const Fr* src, size_t n, Fr* partials,
                                                         unsigned* ticket, Fr* d_evals, Fr* h_evals) {
  __shared__ Fr smem[32];
  Fr acc[1] = {Fr::zero()};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    acc[0] = fr_add(acc[0], ld_fr(src + i));
  grid_finish<1>(acc, smem, partials, ticket, d_evals, h_evals);
}

// rows x cols matrix of Fr -> column sums (multi-GPU fold of gathered partial evaluations)
__global__ void sum_rows_kernel(const Fr* rows, size_t n_rows, size_t cols, Fr* d_out, Fr* h_out) {
  const size_t c = threadIdx.x;
  if (c >= cols) return;
  Fr acc = Fr::zero();
  for (size_t r = 0; r < n_rows; r++) acc = fr_add(acc, ld_fr(rows + r * cols + c));
  st_fr(d_out + c, acc);
  st_fr(h_out + c, acc);
  __threadfence_system();
}

// EqPolynomial.evalsSliceWithScaling (src/poly/mod.zig:252-290), one doubling level per launch:
// for i < size: y = res[i]*r_j; res[i+size] = y; res[i] -= y.
__global__ void __launch_bounds__(kScThreads) eq_level_kernel(Fr* res, size_t size, Fr rj) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < size; i += (size_t)gridDim.x * blockDim.x) {
    Fr x = res[i];
    Fr y = fr_mul(x, rj);
    st_fr(res + i + size, y);
    st_fr(res + i, fr_sub(x, y));
  }
}

__global__ void set_one_kernel(Fr* p, Fr v) { st_fr(p, v); }

// GruenSplitEqPolynomial prefix tables (src/poly/split_eq.zig:131-171): level k+1 from level k,
// next[2i] = prev[i]*(1-tau_k), next[2i+1] = prev[i]*tau_k.  Tables are stored back to back, level
// L at offset 2^L - 1.  One block walks all levels (tables are ~sqrt(N) long).
__global__ void __launch_bounds__(1024) split_eq_tables_kernel(Fr* tab, const Fr* tau, int levels) {
  if (threadIdx.x == 0) st_fr(tab, Fr::one());
  __syncthreads();
  for (int k = 0; k < levels; k++) {
    const size_t prev_size = size_t(1) << k;
    const Fr* prev = tab + (prev_size - 1);
    Fr* next = tab + (2 * prev_size - 1);
    const Fr tk = tau[k];
    const Fr omt = fr_sub(Fr::one(), tk);
    for (size_t i = threadIdx.x; i < prev_size; i += blockDim.x) {
      Fr p = prev[i];
      st_fr(next + 2 * i, fr_mul(p, omt));
      st_fr(next + 2 * i + 1, fr_mul(p, tk));
    }
    __syncthreads();
  }
}

inline Fr fr_from_host(const uint64_t* p) {
  Fr r;
  for (int i = 0; i < 4; i++) {
    r.v[2 * i] = (uint32_t)p[i];
    r.v[2 * i + 1] = (uint32_t)(p[i] >> 32);
  }
  return r;
}

inline unsigned grid_for(const zolt_ctx* ctx, size_t work_items, int per_sm) {
  size_t blocks = (work_items + kScThreads - 1) / kScThreads;
  size_t cap = (size_t)ctx->num_sms * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// instance
// ------------------------------------------------------------------------------------------------
struct zolt_sc {
  zolt_ctx* ctx = nullptr;
  int combiner = 0, pairing = 0, log_n0 = 0, k = 0, n_extra = 0, ne = 0;
  size_t n = 0;  // current table length
  // Round 0/1 read the input tables (owned uploads or borrowed device pointers, never written);
  // later rounds ping-pong between buf[0] (n0/2 entries) and buf[1] (n0/4 entries).
  const Fr* tab_in[kMaxTables] = {};
  bool owns_in = false;
  Fr* buf[2][kMaxTables] = {};
  size_t buf_bytes[2] = {0, 0}, in_bytes = 0;
  int cur = -1;  // -1: current tables are tab_in; 0/1: buf[cur]
  bool evals_valid = false;
  Fr* d_partials = nullptr;
  unsigned* d_ticket = nullptr;
  Fr* d_evals = nullptr;
  Fr* h_evals = nullptr;   // mapped pinned host memory
  Fr* h_evals_dev = nullptr;
  // split-eq
  Fr* d_e_out = nullptr;
  Fr* d_e_in = nullptr;
  Fr* d_tau = nullptr;
  int n_tau = 0;
  unsigned max_grid = 0;
};

namespace {

int combiner_k(int comb) {
  switch (comb) {
    case ZOLT_SC_SUM: return 1;
    case ZOLT_SC_PROD2: return 2;
    case ZOLT_SC_PROD3: return 3;
    case ZOLT_SC_EQ_A_BMC: return 4;
    case ZOLT_SC_EQ_AB_MINUS_C: return 4;
    case ZOLT_SC_GRUEN_AB: return 2;
    case ZOLT_SC_GRUEN_AB_MINUS_C: return 3;
    default: return -1;
  }
}
int combiner_ne(int comb) {
  return (comb == ZOLT_SC_SUM || comb == ZOLT_SC_GRUEN_AB || comb == ZOLT_SC_GRUEN_AB_MINUS_C) ? 2 : 4;
}
bool combiner_gruen(int comb) { return comb == ZOLT_SC_GRUEN_AB || comb == ZOLT_SC_GRUEN_AB_MINUS_C; }

SplitEqDev split_eq_for(const zolt_sc* sc, size_t n_cur) {
  // getWindowEqTables with window_size 1 (src/poly/split_eq.zig:311-343); u = unbound variables.
  SplitEqDev s{nullptr, nullptr, 0};
  if (!combiner_gruen(sc->combiner)) return s;
  int u = 0;
  while ((size_t(1) << u) < n_cur) u++;
  const int head_len = u > 0 ? u - 1 : 0;
  const int m = sc->n_tau / 2;
  const int out_bits = head_len < m ? head_len : m;
  const int in_bits = head_len - out_bits;
  s.e_out = sc->d_e_out + ((size_t(1) << out_bits) - 1);
  s.e_in = sc->d_e_in + ((size_t(1) << in_bits) - 1);
  s.in_bits = in_bits;
  return s;
}

template <int COMB, int PAIR>
void launch_round(zolt_sc* sc, bool fused, const Fr& r) {
  ScPtrs t{};
  const int total = sc->k + sc->n_extra;
  for (int i = 0; i < total; i++) {
    t.src[i] = sc->cur < 0 ? sc->tab_in[i] : sc->buf[sc->cur][i];
    t.dst[i] = sc->buf[sc->cur < 0 ? 0 : (sc->cur ^ 1)][i];
  }
  const size_t n_cur = fused ? sc->n / 2 : sc->n;
  const unsigned grid = grid_for(sc->ctx, n_cur / 2, 4);
  SplitEqDev sq = split_eq_for(sc, n_cur);
  const int ps = prof_begin(sc->ctx, fused ? PROF_SC_FUSED : PROF_SC_EVAL);
  if (fused)
    sc_round_kernel<COMB, PAIR, true><<<grid, kScThreads, 0, sc->ctx->stream>>>(
        t, sc->n_extra, sc->n, r, sq, sc->d_partials, sc->d_ticket, sc->d_evals, sc->h_evals_dev);
  else
    sc_round_kernel<COMB, PAIR, false><<<grid, kScThreads, 0, sc->ctx->stream>>>(
        t, sc->n_extra, sc->n, r, sq, sc->d_partials, sc->d_ticket, sc->d_evals, sc->h_evals_dev);
  prof_end(sc->ctx, ps);
  sc->ctx->launches++;
}

template <int COMB>
void launch_round_p(zolt_sc* sc, bool fused, const Fr& r) {
  if (sc->pairing == ZOLT_PAIR_ADJACENT) launch_round<COMB, ZOLT_PAIR_ADJACENT>(sc, fused, r);
  else launch_round<COMB, ZOLT_PAIR_HALF>(sc, fused, r);
}

void launch_round_any(zolt_sc* sc, bool fused, const Fr& r) {
  switch (sc->combiner) {
    case ZOLT_SC_SUM: launch_round_p<ZOLT_SC_SUM>(sc, fused, r); break;
    case ZOLT_SC_PROD2: launch_round_p<ZOLT_SC_PROD2>(sc, fused, r); break;
    case ZOLT_SC_PROD3: launch_round_p<ZOLT_SC_PROD3>(sc, fused, r); break;
    case ZOLT_SC_EQ_A_BMC: launch_round_p<ZOLT_SC_EQ_A_BMC>(sc, fused, r); break;
    case ZOLT_SC_EQ_AB_MINUS_C: launch_round_p<ZOLT_SC_EQ_AB_MINUS_C>(sc, fused, r); break;
    case ZOLT_SC_GRUEN_AB: launch_round_p<ZOLT_SC_GRUEN_AB>(sc, fused, r); break;
    default: launch_round_p<ZOLT_SC_GRUEN_AB_MINUS_C>(sc, fused, r); break;
  }
}

int sc_alloc(zolt_ctx* ctx, int combiner, int pairing, int log_n, int n_tables, int n_extra, zolt_sc** out) {
  if (!ctx || !out) return ZOLT_ERR_ARG;
  *out = nullptr;
  const int k = combiner_k(combiner);
  if (k < 0 || n_tables != k) return ZOLT_ERR_ARG;
  if (pairing != ZOLT_PAIR_ADJACENT && pairing != ZOLT_PAIR_HALF) return ZOLT_ERR_ARG;
  if (combiner_gruen(combiner) && pairing != ZOLT_PAIR_ADJACENT) return ZOLT_ERR_ARG;  // Gruen provers bind low-to-high only
  if (log_n < 0 || log_n > 40 || n_extra < 0 || k + n_extra > kMaxTables) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  zolt_sc* sc = new (std::nothrow) zolt_sc();
  if (!sc) return ZOLT_ERR_ARG;
  sc->ctx = ctx;
  sc->combiner = combiner;
  sc->pairing = pairing;
  sc->log_n0 = log_n;
  sc->k = k;
  sc->n_extra = n_extra;
  sc->ne = combiner_ne(combiner);
  sc->n = size_t(1) << log_n;
  const size_t n = sc->n;
  sc->buf_bytes[0] = (n / 2 ? n / 2 : 1) * sizeof(Fr);
  sc->buf_bytes[1] = (n / 4 ? n / 4 : 1) * sizeof(Fr);
  sc->in_bytes = n * sizeof(Fr);
  for (int i = 0; i < k + n_extra; i++) {
    ZOLT_CUDA_TRY(pool_alloc(ctx, sc->buf_bytes[0], (void**)&sc->buf[0][i]));
    ZOLT_CUDA_TRY(pool_alloc(ctx, sc->buf_bytes[1], (void**)&sc->buf[1][i]));
  }
  sc->max_grid = (unsigned)ctx->num_sms * 4;
  ZOLT_CUDA_TRY(pool_alloc(ctx, (size_t)sc->max_grid * 4 * sizeof(Fr), (void**)&sc->d_partials));
  ZOLT_CUDA_TRY(pool_alloc(ctx, 256, (void**)&sc->d_ticket));
  ZOLT_CUDA_TRY(cudaMemsetAsync(sc->d_ticket, 0, sizeof(unsigned), ctx->stream));
  ZOLT_CUDA_TRY(pool_alloc(ctx, 4 * sizeof(Fr), (void**)&sc->d_evals));
  ZOLT_CUDA_TRY(pinned_mailbox_alloc(ctx, (void**)&sc->h_evals, (void**)&sc->h_evals_dev));
  *out = sc;
  return ZOLT_OK;
}

}  // namespace

extern "C" int zolt_sc_create(zolt_ctx* ctx, int combiner, int pairing, int log_n, int n_tables, int n_extra,
                              const uint64_t* const* tables, zolt_sc** out) {
  if (!tables) return ZOLT_ERR_ARG;
  zolt_sc* sc = nullptr;
  ZOLT_TRY(sc_alloc(ctx, combiner, pairing, log_n, n_tables, n_extra, &sc));
  sc->owns_in = true;
  for (int i = 0; i < n_tables + n_extra; i++) {
    if (!tables[i]) { zolt_sc_free(sc); return ZOLT_ERR_ARG; }
    Fr* d = nullptr;
    ZOLT_CUDA_TRY(pool_alloc(ctx, sc->in_bytes, (void**)&d));
    sc->tab_in[i] = d;
    ZOLT_CUDA_TRY(cudaMemcpyAsync(d, tables[i], sc->n * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
  }
  ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // host buffers are only borrowed for the call
  *out = sc;
  return ZOLT_OK;
}

extern "C" int zolt_sc_create_dev(zolt_ctx* ctx, int combiner, int pairing, int log_n, int n_tables, int n_extra,
                                  const void* const* d_tables, zolt_sc** out) {
  if (!d_tables) return ZOLT_ERR_ARG;
  zolt_sc* sc = nullptr;
  ZOLT_TRY(sc_alloc(ctx, combiner, pairing, log_n, n_tables, n_extra, &sc));
  sc->owns_in = false;
  for (int i = 0; i < n_tables + n_extra; i++) {
    if (!d_tables[i]) { zolt_sc_free(sc); return ZOLT_ERR_ARG; }
    sc->tab_in[i] = static_cast<const Fr*>(d_tables[i]);  // borrowed, read-only
  }
  *out = sc;
  return ZOLT_OK;
}

extern "C" int zolt_sc_set_split_eq(zolt_sc* sc, const uint64_t* tau, int n_tau) {
  if (!sc || !tau || !combiner_gruen(sc->combiner)) return ZOLT_ERR_ARG;
  if (n_tau != sc->log_n0 || n_tau < 1) return ZOLT_ERR_ARG;
  zolt_ctx* ctx = sc->ctx;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  // w_out = tau[0..m], w_in = tau[m..n-1], w_last skipped (src/poly/split_eq.zig:63-110)
  const int m = n_tau / 2;
  const int n_in = n_tau - 1 - m;
  sc->n_tau = n_tau;
  ZOLT_CUDA_TRY(cudaMalloc((void**)&sc->d_tau, (size_t)n_tau * sizeof(Fr)));
  ZOLT_CUDA_TRY(cudaMemcpyAsync(sc->d_tau, tau, (size_t)n_tau * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
  ZOLT_CUDA_TRY(cudaMalloc((void**)&sc->d_e_out, ((size_t(2) << m) - 1) * sizeof(Fr)));
  ZOLT_CUDA_TRY(cudaMalloc((void**)&sc->d_e_in, ((size_t(2) << n_in) - 1) * sizeof(Fr)));
  split_eq_tables_kernel<<<1, 1024, 0, ctx->stream>>>(sc->d_e_out, sc->d_tau, m);
  split_eq_tables_kernel<<<1, 1024, 0, ctx->stream>>>(sc->d_e_in, sc->d_tau + m, n_in);
  ctx->launches += 2;
  ZOLT_CUDA_TRY(cudaGetLastError());
  ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // tau is a borrowed host buffer
  return ZOLT_OK;
}

extern "C" int zolt_sc_num_evals(const zolt_sc* sc, int* out) {
  if (!sc || !out) return ZOLT_ERR_ARG;
  *out = sc->ne;
  return ZOLT_OK;
}

static int sc_ensure_evals(zolt_sc* sc) {
  if (sc->evals_valid) return ZOLT_OK;
  if (sc->n < 2) return ZOLT_ERR_STATE;  // all variables bound: no round polynomial left
  if (combiner_gruen(sc->combiner) && !sc->d_e_out) return ZOLT_ERR_STATE;
  ZOLT_CUDA_TRY(cudaSetDevice(sc->ctx->device));
  launch_round_any(sc, false, Fr::zero());
  ZOLT_CUDA_TRY(cudaGetLastError());
  sc->evals_valid = true;
  return ZOLT_OK;
}

extern "C" int zolt_sc_round_evals(zolt_sc* sc, uint64_t* evals_out) {
  if (!sc || !evals_out) return ZOLT_ERR_ARG;
  ZOLT_TRY(sc_ensure_evals(sc));
  ZOLT_CUDA_TRY(cudaStreamSynchronize(sc->ctx->stream));
  memcpy(evals_out, sc->h_evals, (size_t)sc->ne * sizeof(Fr));
  return ZOLT_OK;
}

extern "C" int zolt_sc_round_partials_dev(zolt_sc* sc, const void** d_ptr) {
  if (!sc || !d_ptr) return ZOLT_ERR_ARG;
  ZOLT_TRY(sc_ensure_evals(sc));
  *d_ptr = sc->d_evals;
  return ZOLT_OK;
}

extern "C" int zolt_sc_bind(zolt_sc* sc, const uint64_t challenge[4]) {
  if (!sc || !challenge) return ZOLT_ERR_ARG;
  if (sc->n < 2) return ZOLT_ERR_STATE;
  if (combiner_gruen(sc->combiner) && !sc->d_e_out) return ZOLT_ERR_STATE;
  zolt_ctx* ctx = sc->ctx;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  const Fr r = fr_from_host(challenge);
  if (sc->n >= 4) {
    launch_round_any(sc, true, r);
    sc->evals_valid = true;
  } else {
    // last variable: bind only, nothing left to evaluate
    for (int i = 0; i < sc->k + sc->n_extra; i++) {
      const Fr* src = sc->cur < 0 ? sc->tab_in[i] : sc->buf[sc->cur][i];
      Fr* dst = sc->buf[sc->cur < 0 ? 0 : (sc->cur ^ 1)][i];
      if (sc->pairing == ZOLT_PAIR_ADJACENT)
        bind_kernel<ZOLT_PAIR_ADJACENT><<<1, kScThreads, 0, ctx->stream>>>(dst, src, sc->n, r);
      else
        bind_kernel<ZOLT_PAIR_HALF><<<1, kScThreads, 0, ctx->stream>>>(dst, src, sc->n, r);
      ctx->launches++;
    }
    sc->evals_valid = false;
  }
  ZOLT_CUDA_TRY(cudaGetLastError());
  sc->cur = sc->cur < 0 ? 0 : (sc->cur ^ 1);
  sc->n /= 2;
  return ZOLT_OK;
}

extern "C" int zolt_sc_len(const zolt_sc* sc, size_t* n) {
  if (!sc || !n) return ZOLT_ERR_ARG;
  *n = sc->n;
  return ZOLT_OK;
}

extern "C" int zolt_sc_download(zolt_sc* sc, int table, uint64_t* dst) {
  if (!sc || !dst || table < 0 || table >= sc->k + sc->n_extra) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(sc->ctx->device));
  const Fr* src = sc->cur < 0 ? sc->tab_in[table] : sc->buf[sc->cur][table];
  ZOLT_CUDA_TRY(cudaMemcpyAsync(dst, src, sc->n * sizeof(Fr), cudaMemcpyDeviceToHost, sc->ctx->stream));
  ZOLT_CUDA_TRY(cudaStreamSynchronize(sc->ctx->stream));
  return ZOLT_OK;
}

extern "C" int zolt_sc_table_dev(zolt_sc* sc, int table, const void** d_ptr) {
  if (!sc || !d_ptr || table < 0 || table >= sc->k + sc->n_extra) return ZOLT_ERR_ARG;
  *d_ptr = sc->cur < 0 ? sc->tab_in[table] : sc->buf[sc->cur][table];
  return ZOLT_OK;
}

extern "C" int zolt_sc_free(zolt_sc* sc) {
  if (!sc) return ZOLT_ERR_ARG;
  cudaSetDevice(sc->ctx->device);
  cudaStreamSynchronize(sc->ctx->stream);
  zolt_ctx* ctx = sc->ctx;
  for (int i = 0; i < kMaxTables; i++) {
    if (sc->owns_in && sc->tab_in[i]) pool_release(ctx, const_cast<Fr*>(sc->tab_in[i]), sc->in_bytes);
    if (sc->buf[0][i]) pool_release(ctx, sc->buf[0][i], sc->buf_bytes[0]);
    if (sc->buf[1][i]) pool_release(ctx, sc->buf[1][i], sc->buf_bytes[1]);
  }
  if (sc->d_partials) pool_release(ctx, sc->d_partials, (size_t)sc->max_grid * 4 * sizeof(Fr));
  if (sc->d_ticket) pool_release(ctx, sc->d_ticket, 256);
  if (sc->d_evals) pool_release(ctx, sc->d_evals, 4 * sizeof(Fr));
  pinned_mailbox_release(ctx, sc->h_evals);
  if (sc->d_e_out) cudaFree(sc->d_e_out);
  if (sc->d_e_in) cudaFree(sc->d_e_in);
  if (sc->d_tau) cudaFree(sc->d_tau);
  delete sc;
  return ZOLT_OK;
}

// ------------------------------------------------------------------------------------------------
// stand-alone table ops
// ------------------------------------------------------------------------------------------------
namespace {
struct Scratch {
  Fr* partials = nullptr;
  unsigned* ticket = nullptr;
  Fr* d_out = nullptr;
};
// per-context scratch for zolt_poly_sum / zolt_fr_sum_rows, allocated on demand and freed per call
int scratch_alloc(zolt_ctx* ctx, Scratch* s) {
  ZOLT_CUDA_TRY(pool_alloc(ctx, (size_t)ctx->num_sms * 4 * 4 * sizeof(Fr), (void**)&s->partials));
  ZOLT_CUDA_TRY(pool_alloc(ctx, 256, (void**)&s->ticket));
  ZOLT_CUDA_TRY(cudaMemsetAsync(s->ticket, 0, sizeof(unsigned), ctx->stream));
  ZOLT_CUDA_TRY(pool_alloc(ctx, 16 * sizeof(Fr), (void**)&s->d_out));
  return ZOLT_OK;
}
void scratch_free(zolt_ctx* ctx, Scratch* s) {
  pool_release(ctx, s->partials, (size_t)ctx->num_sms * 4 * 4 * sizeof(Fr));
  pool_release(ctx, s->ticket, 256);
  pool_release(ctx, s->d_out, 16 * sizeof(Fr));
}
}  // namespace

extern "C" int zolt_poly_bind_low(zolt_ctx* ctx, void* d_dst, const void* d_src, size_t n, const uint64_t r[4]) {
  if (!ctx || !d_dst || !d_src || !r || n < 2 || (n & (n - 1))) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  const Fr rr = fr_from_host(r);
  const unsigned grid = grid_for(ctx, n / 2, 8);
  if (d_dst == d_src) {
    // in place (the reference's contract): thread i would overwrite element i while thread i/2 may
    // still need it, so bind into a scratch buffer and copy back on the same stream.
    Fr* tmp = nullptr;
    ZOLT_CUDA_TRY(cudaMalloc((void**)&tmp, (n / 2) * sizeof(Fr)));
    bind_kernel<ZOLT_PAIR_ADJACENT><<<grid, kScThreads, 0, ctx->stream>>>(tmp, (const Fr*)d_src, n, rr);
    ctx->launches++;
    ZOLT_CUDA_TRY(cudaGetLastError());
    ZOLT_CUDA_TRY(cudaMemcpyAsync(d_dst, tmp, (n / 2) * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
    ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    cudaFree(tmp);
  } else {
    bind_kernel<ZOLT_PAIR_ADJACENT><<<grid, kScThreads, 0, ctx->stream>>>((Fr*)d_dst, (const Fr*)d_src, n, rr);
    ctx->launches++;
    ZOLT_CUDA_TRY(cudaGetLastError());
  }
  return ZOLT_OK;
}

extern "C" int zolt_poly_bind_first(zolt_ctx* ctx, void* d_dst, const void* d_src, size_t n, const uint64_t r[4]) {
  if (!ctx || !d_dst || !d_src || !r || n < 2 || (n & (n - 1))) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  const Fr rr = fr_from_host(r);
  const unsigned grid = grid_for(ctx, n / 2, 8);
  // half-split: element i only depends on i and i+n/2, so in-place (dst == src) is race-free
  bind_kernel<ZOLT_PAIR_HALF><<<grid, kScThreads, 0, ctx->stream>>>((Fr*)d_dst, (const Fr*)d_src, n, rr);
  ctx->launches++;
  ZOLT_CUDA_TRY(cudaGetLastError());
  return ZOLT_OK;
}

extern "C" int zolt_eq_evals(zolt_ctx* ctx, const uint64_t* r, int n_vars, const uint64_t* scale, void* d_out) {
  if (!ctx || !d_out || n_vars < 0 || n_vars > 40 || (n_vars > 0 && !r)) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  Fr* res = (Fr*)d_out;
  set_one_kernel<<<1, 1, 0, ctx->stream>>>(res, scale ? fr_from_host(scale) : Fr::one());
  ctx->launches++;
  size_t size = 1;
  for (int j = n_vars - 1; j >= 0; j--) {
    eq_level_kernel<<<grid_for(ctx, size, 8), kScThreads, 0, ctx->stream>>>(res, size, fr_from_host(r + 4 * j));
    ctx->launches++;
    size *= 2;
  }
  ZOLT_CUDA_TRY(cudaGetLastError());
  return ZOLT_OK;
}

extern "C" int zolt_poly_sum(zolt_ctx* ctx, const void* d_src, size_t n, uint64_t out[4]) {
  if (!ctx || !d_src || !out) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  Scratch s;
  ZOLT_TRY(scratch_alloc(ctx, &s));
  sum_kernel<<<grid_for(ctx, n, 4), kScThreads, 0, ctx->stream>>>((const Fr*)d_src, n, s.partials, s.ticket, s.d_out,
                                                                  (Fr*)ctx->d_mailbox);
  ctx->launches++;
  ZOLT_CUDA_TRY(cudaGetLastError());
  ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  memcpy(out, ctx->h_mailbox, 32);
  scratch_free(ctx, &s);
  return ZOLT_OK;
}

extern "C" int zolt_fr_sum_rows(zolt_ctx* ctx, const void* d_rows, size_t rows, size_t cols, uint64_t* out) {
  if (!ctx || !d_rows || !out || cols == 0 || cols > 8) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  Scratch s;
  ZOLT_TRY(scratch_alloc(ctx, &s));
  sum_rows_kernel<<<1, 32, 0, ctx->stream>>>((const Fr*)d_rows, rows, cols, s.d_out, (Fr*)ctx->d_mailbox);
  ctx->launches++;
  ZOLT_CUDA_TRY(cudaGetLastError());
  ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  memcpy(out, ctx->h_mailbox, cols * 32);
  scratch_free(ctx, &s);
  return ZOLT_OK;
}
