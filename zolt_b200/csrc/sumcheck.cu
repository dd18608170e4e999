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
#include <cuda.h>            // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cudaTypedefs.h>

#include <emmintrin.h>

#include <atomic>
#include <chrono>
#include <cstring>
#include <new>

#include "common.cuh"
#include "field.cuh"
#include "peer.cuh"

using namespace zolt;

namespace {

constexpr int kMaxTables = 8;
constexpr int kScThreads = 256;

struct ScPtrs {
  const Fr* src[kMaxTables];
  Fr* dst[kMaxTables];
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
// extrapolation to X = 2 (hi + (hi - lo)) feeding a product: stays below 2p, the product's reduction absorbs it
__device__ __forceinline__ Fr fr_add_lazy(const Fr& a, const Fr& b) { return fp_add_lazy<FrParams>(a, b); }
__device__ __forceinline__ Fr fr_mul(const Fr& a, const Fr& b) { return fp_mul<FrParams>(a, b); }

struct SplitEqDev {
  const Fr* e_out;  // level `out_bits` table of E_out_vec
  const Fr* e_in;   // level `in_bits` table of E_in_vec
  int in_bits;
  unsigned out_mask = 0xffffffffu;  // bits of (pair >> in_bits) that index e_out (Stage 4: the row k above them is masked out)
  __device__ __forceinline__ Fr weight(size_t g) const {
    // E_out[x_out] * E_in[x_in], pair g = (x_out << in_bits) | x_in  (src/zkvm/spartan/product_remainder.zig:293-296)
    const size_t x_out = (g >> in_bits) & out_mask, x_in = g & ((size_t(1) << in_bits) - 1);
    return fr_mul(ld_fr(e_out + x_out), ld_fr(e_in + x_in));
  }
};


// lo + r*(hi - lo): DensePolynomial.bindLow's form (src/poly/mod.zig:171); equal, as a canonical
// field element, to the (1-r)*lo + r*hi form used at val_evaluation.zig:626-631.
__device__ __forceinline__ Fr bind_pair(const Fr& lo, const Fr& hi, const Fr& r) {
  return fr_add(lo, fr_mul(r, fr_sub(hi, lo)));
}

// ---- combiners ------------------------------------------------------------------------------------
// Every product combiner keeps its running sums as UNREDUCED 512-bit accumulators (field.cuh wide_*):
// the last multiplication of each term costs a plain product, and the Montgomery reduction happens
// once per thread.  The round polynomial is sampled at X = 0, 2 and "infinity" (its leading
// coefficient); p(1) comes from the sumcheck identity p(0) + p(1) = claim and p(3) from the third
// finite difference, both exact in the field, so the four evaluations handed to the host are the
// canonical values the reference computes by direct summation (e.g. output_check.zig:399-432).
//
// Raw sums leave the kernel as 4 field elements:
//   cubic      [p(0), p(1) or 0, p(2), lead]     quadratic [p(0), p(1) or 0, lead, 0]
//   SUM/Gruen  [s0, s1, 0, 0]                    Stage-4 dense [p(0), p(2), p(3), 0]
template <int COMB> struct CombTraits;
template <> struct CombTraits<ZOLT_SC_SUM> { static constexpr int K = 1, NE = 2, NW = 0, DEG = 0; static constexpr bool GRUEN = false; };
template <> struct CombTraits<ZOLT_SC_PROD2> { static constexpr int K = 2, NE = 4, NW = 2, DEG = 2; static constexpr bool GRUEN = false; };
template <> struct CombTraits<ZOLT_SC_PROD3> { static constexpr int K = 3, NE = 4, NW = 3, DEG = 3; static constexpr bool GRUEN = false; };
template <> struct CombTraits<ZOLT_SC_EQ_A_BMC> { static constexpr int K = 4, NE = 4, NW = 3, DEG = 3; static constexpr bool GRUEN = false; };
template <> struct CombTraits<ZOLT_SC_EQ_AB_MINUS_C> { static constexpr int K = 4, NE = 4, NW = 3, DEG = 3; static constexpr bool GRUEN = false; };
template <> struct CombTraits<ZOLT_SC_GRUEN_AB> { static constexpr int K = 2, NE = 2, NW = 2, DEG = 0; static constexpr bool GRUEN = true; };
template <> struct CombTraits<ZOLT_SC_GRUEN_AB_MINUS_C> { static constexpr int K = 3, NE = 2, NW = 2, DEG = 0; static constexpr bool GRUEN = true; };
template <> struct CombTraits<ZOLT_SC_STAGE4_GRUEN> { static constexpr int K = 4, NE = 2, NW = 2, DEG = 0; static constexpr bool GRUEN = true; };
template <> struct CombTraits<ZOLT_SC_STAGE4_DENSE> { static constexpr int K = 5, NE = 3, NW = 3, DEG = 0; static constexpr bool GRUEN = false; };
template <> struct CombTraits<ZOLT_SC_EQ_DOT2> { static constexpr int K = 5, NE = 3, NW = 3, DEG = 0; static constexpr bool GRUEN = false; };

constexpr int kRaw = 4;  // field elements per partial result

// ---- per-thread shared-memory scratch --------------------------------------------------------------
// Every thread of a round kernel owns 512 bytes of shared memory, addressed in 16-byte units laid out
// [unit][thread] so that a warp's 128-bit accesses are conflict-free.  It is PRIVATE to the thread (no
// barrier is ever needed) and serves as
//   * the landing zone of cp.async prefetches: the source elements a thread needs next are copied
//     global -> shared while it is still multiplying, so no register is spent on loads in flight and
//     the HBM latency is off the critical path;
//   * the stash of the bound (lo, hi) values of every table between the bind loop, which is rolled over
//     the tables to keep the kernel's code small enough for the instruction cache, and the evaluation.
constexpr int kRoundThreads = 128;
constexpr int kEvalBlocksPerSm = 3;   // MODE 0/1: 32 scratch units (512 B) per thread
constexpr int kFusedBlocksPerSm = 4;  // MODE 2: 24 scratch units (384 B) per thread
constexpr int kEvalScratchUnits = 32;
constexpr int kFusedScratchUnits = 24;

// Unit stride of the scratch, in 16-byte slots: the thread count padded so that it is 4 mod 8.  A thread's own accesses
// (32 consecutive slots per warp) are conflict-free for any stride; the padding makes the warp-cooperative copies
// below conflict-free too (their 32 destinations spread over (unit, thread) instead of thread alone).
constexpr int kScrStride = kRoundThreads + 4;

__device__ __forceinline__ Fr scr_ld(const uint4* base, int unit) {
  const uint4 a = base[unit * kScrStride], b = base[(unit + 1) * kScrStride];
  Fr r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void scr_st(uint4* base, int unit, const Fr& x) {
  base[unit * kScrStride] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
  base[(unit + 1) * kScrStride] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}
// ---- coalesced global -> scratch copies ----------------------------------------------------------------------------
// A thread's source elements are 32-byte values 64 or 128 bytes apart from its neighbour's: copied by their owner, one
// warp instruction touches 32 different 128-byte lines, 16 bytes each, and every 32-byte sector crosses L2 -> SM twice
// (r2 v4 ablation: the loads ALONE ran at 3.4 TB/s and bounded every round, whatever the arithmetic did).  Instead the
// warp copies its contiguous source range cooperatively: instruction j moves the 512 contiguous bytes
// [512 j, 512 j + 512), lane L the 16-byte piece q = 32 j + L, and drops it into the scratch slot of the thread that
// owns it.  `pieces_per_thread` consecutive pieces belong to one thread and land in its units unit0, unit0 + 1, ...
// (piece order == unit order: an element's low half, its high half, the next element ...).  Pieces at or beyond
// `valid` (the ragged end of the table) are skipped.  Every lane of the warp must call this (converged).
template <int PIECES_PER_THREAD>
__device__ __forceinline__ void warp_copy_run(uint4* warp_scr, int unit0, const Fr* src, int valid, int lane) {
  constexpr int kInstr = PIECES_PER_THREAD;  // 32 threads x PIECES_PER_THREAD pieces / 32 lanes
  const uint4* g = reinterpret_cast<const uint4*>(src);
#pragma unroll
  for (int j = 0; j < kInstr; j++) {
    const int q = 32 * j + lane;
    const int owner = q / PIECES_PER_THREAD, unit = unit0 + q % PIECES_PER_THREAD;
    if (q < valid) {
      const unsigned dst = (unsigned)__cvta_generic_to_shared(warp_scr + unit * kScrStride + owner);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(g + q) : "memory");
    }
  }
}
// ---- TMA global -> shared copies -------------------------------------------------------------------------------------
// cp.async (LDGSTS) writes shared memory 16 bytes per cycle per SM on this part: 148 x 16 B x 1.965 GHz = 4.6 TB/s is all
// a kernel can stream through it, well below HBM (r2 v5: the loads ALONE ran at 3.5-3.8 TB/s, coalesced or not, with
// 32 shared-memory wavefronts per copy instruction).  The adjacent-pair rounds therefore move their sources with TMA
// tensor copies: a table is described to the TMA unit as a 2-D tensor of ROWS (one row = everything one pair reads
// from it: 128 bytes in a fused pass, 64 in an evaluation-only round), a warp's 32 consecutive pairs are one box of 32
// rows, fetched by ONE instruction of one lane and counted in bytes on a per-warp mbarrier.  The box lands in shared
// memory with the hardware swizzle of its row width (16-byte chunk index XOR row bits), which is exactly what makes the
// 8 lanes served together by a 128-bit shared load hit 8 different bank groups although their rows are one bank period
// apart.  Rows past the end of the table are zero-filled by the TMA unit (and never used).  Measured (r2 v8, 2^24 entries,
// 4 tables): loads alone 6.6 TB/s, loads + bound-table stores 6.0 TB/s -- the memory side of a pass is at the HBM roofline;
// what is left is the arithmetic.
//   r2 v6/v7 (1-D bulk copies, 8 per table and warp into padded pieces): same bandwidth, but the issue sequence
//   (5 R2UR + UBLKCP per piece, serialised) cost 20 % of the pass; the tensor form is one instruction.
struct alignas(64) TmapArray {
  CUtensorMap m[kMaxTables];
};
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// box of 32 rows starting at row `row0` of the tensor -> dst (swizzled); completion bytes on `bar`.  One lane.
__device__ __forceinline__ void tma_rows_g2s(void* dst, const CUtensorMap* tm, int row0, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
               "l"(tm), "r"(0), "r"(row0), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// acc += a * b (unreduced), then restore acc < 2^511
__device__ __forceinline__ void wide_acc(uint32_t* acc, const Fr& a, const Fr& b) {
  wide_mul_add(acc, a.v, b.v);
  wide_normalize<FrParams>(acc);
}

// One pair of every combiner table; the values sit in the thread's stash at units 4k (lo) and 4k+2 (hi).
// wacc: NW (+1 when P1) wide accumulators, facc: SUM only.
// Where a thread finds the (lo, hi) values of table k:
//   UnitStash  [unit][thread] slots of 16 bytes (unit 4k = lo, 4k+2 = hi), thread stride `S` slots
//   SwizzledRows64  one 64-byte row (lo, hi) per table inside the swizzled TMA boxes of the evaluation-only rounds
template <int S>
struct UnitStash {
  const uint4* base;
  __device__ __forceinline__ Fr ld(int unit) const {
    const uint4 a = base[unit * S], b = base[(unit + 1) * S];
    Fr r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
  }
  __device__ __forceinline__ Fr lo(int k) const { return ld(4 * k); }
  __device__ __forceinline__ Fr hi(int k) const { return ld(4 * k + 2); }
};
struct SwizzledRows64 {
  const unsigned char* row;  // this thread's 64-byte row (lo, hi) of table 0, inside a 64-byte-swizzled TMA box
  int sw, pitch;             // chunk c sits at byte (16 c) ^ sw of the row; `pitch` bytes between the boxes of two tables
  __device__ __forceinline__ Fr ld(int k, int c) const {
    const uint4 a = *reinterpret_cast<const uint4*>(row + k * pitch + ((16 * c) ^ sw));
    const uint4 b = *reinterpret_cast<const uint4*>(row + k * pitch + ((16 * (c + 1)) ^ sw));
    Fr r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
  }
  __device__ __forceinline__ Fr lo(int k) const { return ld(k, 0); }
  __device__ __forceinline__ Fr hi(int k) const { return ld(k, 2); }
};

template <int COMB, bool P1, class Stash>
__device__ __forceinline__ void accumulate_pair(uint32_t (*wacc)[16], Fr* facc, const Stash& st, const Fr& w) {
#define LO(k) st.lo(k)
#define HI(k) st.hi(k)
  if constexpr (COMB == ZOLT_SC_SUM) {
    facc[0] = fr_add(facc[0], LO(0));
    facc[1] = fr_add(facc[1], HI(0));
  } else if constexpr (COMB == ZOLT_SC_PROD2) {
    // a(X) b(X): p(0) = a0 b0, lead = (a1-a0)(b1-b0)
    const Fr a0 = LO(0), b0 = LO(1), a1 = HI(0), b1 = HI(1);
    wide_acc(wacc[0], a0, b0);
    if constexpr (P1) wide_acc(wacc[2], a1, b1);
    wide_acc(wacc[1], fr_sub(a1, a0), fr_sub(b1, b0));
  } else if constexpr (COMB == ZOLT_SC_PROD3 || COMB == ZOLT_SC_EQ_A_BMC) {
    // x(X) y(X) z(X);  eq*a*(b-c) is the same shape with z = b - c
    Fr m0, m2, ml, m1;
    {
      const Fr x0 = LO(0), y0 = LO(1), x1 = HI(0), y1 = HI(1);
      m0 = fr_mul(x0, y0);
      const Fr dx = fr_sub(x1, x0), dy = fr_sub(y1, y0);
      ml = fr_mul(dx, dy);
      if constexpr (P1) {
        // with x1 y1 at hand the product at X = 2 needs no multiplication: (2x1-x0)(2y1-y0) = 2 dx dy + 2 x1 y1 - x0 y0
        m1 = fr_mul(x1, y1);
        const Fr s = fr_add(ml, m1);
        m2 = fr_sub(fr_add(s, s), m0);
      } else {
        m2 = fr_mul(fr_add_lazy(x1, dx), fr_add_lazy(y1, dy));  // (< 2p) * (< 2p) < p * 2^256
      }
    }
    Fr z0, z1;
    if constexpr (COMB == ZOLT_SC_PROD3) {
      z0 = LO(2);
      z1 = HI(2);
    } else {
      z0 = fr_sub(LO(2), LO(3));
      z1 = fr_sub(HI(2), HI(3));
    }
    wide_acc(wacc[0], m0, z0);
    if constexpr (P1) wide_acc(wacc[3], m1, z1);
    const Fr dz = fr_sub(z1, z0);
    wide_acc(wacc[2], ml, dz);
    wide_acc(wacc[1], m2, fr_add_lazy(z1, dz));
  } else if constexpr (COMB == ZOLT_SC_EQ_AB_MINUS_C) {
    // e(X) * (a(X) b(X) - c(X)): the leading coefficient has no c term
    Fr t0, t2, ml, t1;
    {
      const Fr a0 = LO(1), b0 = LO(2), a1 = HI(1), b1 = HI(2);
      t0 = fr_mul(a0, b0);
      const Fr da = fr_sub(a1, a0), db = fr_sub(b1, b0);
      ml = fr_mul(da, db);
      if constexpr (P1) {
        // (2a1-a0)(2b1-b0) = 2 da db + 2 a1 b1 - a0 b0: no fourth multiplication when a1 b1 is needed anyway
        t1 = fr_mul(a1, b1);
        const Fr s = fr_add(ml, t1);
        t2 = fr_sub(fr_add(s, s), t0);
      } else {
        t2 = fr_mul(fr_add_lazy(a1, da), fr_add_lazy(b1, db));  // (< 2p) * (< 2p) < p * 2^256
      }
    }
    {
      const Fr c0 = LO(3), c1 = HI(3);
      t0 = fr_sub(t0, c0);
      if constexpr (P1) t1 = fr_sub(t1, c1);
      t2 = fr_sub(t2, fr_add(c1, fr_sub(c1, c0)));
    }
    const Fr e0 = LO(0), e1 = HI(0);
    wide_acc(wacc[0], e0, t0);
    if constexpr (P1) wide_acc(wacc[3], e1, t1);
    const Fr de = fr_sub(e1, e0);
    wide_acc(wacc[2], de, ml);
    wide_acc(wacc[1], fr_add_lazy(e1, de), t2);
  } else if constexpr (COMB == ZOLT_SC_GRUEN_AB || COMB == ZOLT_SC_GRUEN_AB_MINUS_C) {
    const Fr l0 = LO(0), r0 = LO(1), l1 = HI(0), r1 = HI(1);
    Fr p0 = fr_mul(l0, r0);
    if constexpr (COMB == ZOLT_SC_GRUEN_AB_MINUS_C) p0 = fr_sub(p0, LO(2));
    wide_acc(wacc[0], w, p0);
    wide_acc(wacc[1], w, fr_mul(fr_sub(l1, l0), fr_sub(r1, r0)));
  } else if constexpr (COMB == ZOLT_SC_STAGE4_GRUEN) {
    // C(X) = ra val + wa (val + inc): C(0) and its X^2 coefficient, weighted (stage4_gruen_prover.zig:629-700).  Each is two
    // products under one Montgomery reduction.
    Fr c0, cx;
    {
      const Fr ra0 = LO(0), wa0 = LO(1), v0 = LO(2), i0 = LO(3);
      c0 = fp_mul2_add<FrParams>(ra0, v0, wa0, fr_add(v0, i0));
      const Fr dv = fr_sub(HI(2), v0);
      const Fr dvi = fr_add(dv, fr_sub(HI(3), i0));
      cx = fp_mul2_add<FrParams>(fr_sub(HI(0), ra0), dv, fr_sub(HI(1), wa0), dvi);
    }
    wide_acc(wacc[0], w, c0);
    wide_acc(wacc[1], w, cx);
  } else if constexpr (COMB == ZOLT_SC_STAGE4_DENSE) {
    // eq(X) * C(X) at X = 0, 2, 3, the reference's sample points (phase2ComputeMessage / phase3ComputeMessage, :764-1015).
    // p(3) is summed like the reference does rather than derived from the leading coefficient: the caller's claim may be
    // a batched one (prove, :433-536), and then only directly summed values match
    Fr c0, c2, c3;
    {
      const Fr ra0 = LO(1), wa0 = LO(2), v0 = LO(3), i0 = LO(4);
      const Fr ra1 = HI(1), wa1 = HI(2), v1 = HI(3), i1 = HI(4);
      c0 = fp_mul2_add<FrParams>(ra0, v0, wa0, fr_add(v0, i0));
      const Fr dra = fr_sub(ra1, ra0), dwa = fr_sub(wa1, wa0), dv = fr_sub(v1, v0), di = fr_sub(i1, i0);
      const Fr ra2 = fr_add(ra1, dra), wa2 = fr_add(wa1, dwa), v2 = fr_add(v1, dv), i2 = fr_add(i1, di);
      c2 = fp_mul2_add<FrParams>(ra2, v2, wa2, fr_add(v2, i2));
      const Fr v3 = fr_add(v2, dv);
      c3 = fp_mul2_add<FrParams>(fr_add(ra2, dra), v3, fr_add(wa2, dwa), fr_add(v3, fr_add(i2, di)));
    }
    const Fr e0 = LO(0), e1 = HI(0);
    const Fr de = fr_sub(e1, e0);
    const Fr e2 = fr_add(e1, de);
    wide_acc(wacc[0], e0, c0);
    wide_acc(wacc[1], e2, c2);
    wide_acc(wacc[2], fr_add(e2, de), c3);
  } else if constexpr (COMB == ZOLT_SC_EQ_DOT2) {
    // w(X) * (a(X) b(X) + c(X) d(X)) at X = 0, 2, 3: one side of InstructionInputProver.computeRoundEvals
    // (stage3_prover.zig:2029-2105: flag * value + flag * value under the eq weight, sampled where the reference samples).
    // Each inner sum is two products under one Montgomery reduction.
    Fr s0, s2, s3;
    {
      const Fr a0 = LO(1), b0 = LO(2), c0 = LO(3), d0 = LO(4);
      s0 = fp_mul2_add<FrParams>(a0, b0, c0, d0);
      const Fr da = fr_sub(HI(1), a0), db = fr_sub(HI(2), b0), dc = fr_sub(HI(3), c0), dd = fr_sub(HI(4), d0);
      const Fr a2 = fr_add(HI(1), da), b2 = fr_add(HI(2), db), c2 = fr_add(HI(3), dc), d2 = fr_add(HI(4), dd);
      s2 = fp_mul2_add<FrParams>(a2, b2, c2, d2);
      s3 = fp_mul2_add<FrParams>(fr_add(a2, da), fr_add(b2, db), fr_add(c2, dc), fr_add(d2, dd));
    }
    const Fr e0 = LO(0), e1 = HI(0);
    const Fr de = fr_sub(e1, e0);
    const Fr e2 = fr_add(e1, de);
    wide_acc(wacc[0], e0, s0);
    wide_acc(wacc[1], e2, s2);
    wide_acc(wacc[2], fr_add(e2, de), s3);
  }
#undef LO
#undef HI
}

// ---- cold-path helpers ---------------------------------------------------------------------------------
// Everything that runs once per thread or once per launch is kept OUT OF LINE with rolled loops.  Inlined,
// the epilogue of a round kernel was ~60 KB of straight-line code that a single warp stepped through once:
// an instruction-cache miss every 128 bytes, ~40 us per launch whatever the table size (ncu: r1 v3 launch
// list, the 2-entry rounds took as long as the 2^13-entry ones).  One copy of each primitive stays cached.
struct Wide { uint32_t v[16]; };
__device__ __noinline__ Fr wide_reduce_cold(Wide w) { return wide_reduce<FrParams>(w.v); }
__device__ __noinline__ Fr fr_add_cold(Fr a, Fr b) { return fr_add(a, b); }
__device__ __noinline__ Fr fr_sub_cold(Fr a, Fr b) { return fr_sub(a, b); }
__device__ __noinline__ Fr fr_mul_cold(Fr a, Fr b) { return fr_mul(a, b); }

// thread-local accumulators -> the 4 raw sums
__device__ __forceinline__ Fr wide_collapse(const uint32_t* acc) {
  Wide w;
#pragma unroll
  for (int i = 0; i < 16; i++) w.v[i] = acc[i];
  return wide_reduce_cold(w);
}

template <int COMB, bool P1>
__device__ __forceinline__ void collapse(uint32_t (*wacc)[16], const Fr* facc, Fr* raw) {
  constexpr int NW = CombTraits<COMB>::NW;
  constexpr int DEG = CombTraits<COMB>::DEG;
#pragma unroll
  for (int i = 0; i < kRaw; i++) raw[i] = Fr::zero();
  if constexpr (COMB == ZOLT_SC_SUM) {
    raw[0] = facc[0];
    raw[1] = facc[1];
  } else if constexpr (DEG == 3) {
    raw[0] = wide_collapse(wacc[0]);
    raw[2] = wide_collapse(wacc[1]);
    raw[3] = wide_collapse(wacc[2]);
    if constexpr (P1) raw[1] = wide_collapse(wacc[NW]);
  } else if constexpr (DEG == 2) {
    raw[0] = wide_collapse(wacc[0]);
    raw[2] = wide_collapse(wacc[1]);
    if constexpr (P1) raw[1] = wide_collapse(wacc[NW]);
  } else {
    raw[0] = wide_collapse(wacc[0]);
    raw[1] = wide_collapse(wacc[1]);
    if constexpr (NW == 3) raw[2] = wide_collapse(wacc[2]);
  }
}

// ---- raw sums -> the evaluations at X = 0..3 -------------------------------------------------------
enum { CLAIM_NONE = 0 /* p(1) was summed directly */, CLAIM_PREV = 1 /* claim = p_prev(r) */, CLAIM_HOST = 2 /* *d_claim */ };

// p(r) for the cubic through (0,p0) (1,p1) (2,p2) (3,p3): Newton forward differences
__device__ __noinline__ Fr newton_eval4(const Fr* p, const Fr& r) {
  const Fr d1 = fr_sub_cold(p[1], p[0]);
  const Fr d1b = fr_sub_cold(p[2], p[1]);
  const Fr d1c = fr_sub_cold(p[3], p[2]);
  const Fr d2 = fr_sub_cold(d1b, d1);
  const Fr d2b = fr_sub_cold(d1c, d1b);
  const Fr d3 = fr_sub_cold(d2b, d2);
  Fr inv2, inv6;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    inv2.v[i] = FrParams::inv2(i);
    inv6.v[i] = FrParams::inv6(i);
  }
  const Fr rm1 = fr_sub_cold(r, Fr::one());
  const Fr rm2 = fr_sub_cold(rm1, Fr::one());
  const Fr rr1 = fr_mul_cold(r, rm1);
  const Fr c2 = fr_mul_cold(rr1, inv2);                    // r(r-1)/2
  const Fr c3 = fr_mul_cold(fr_mul_cold(rr1, rm2), inv6);  // r(r-1)(r-2)/6
  Fr acc = fr_add_cold(p[0], fr_mul_cold(r, d1));
  acc = fr_add_cold(acc, fr_mul_cold(c2, d2));
  acc = fr_add_cold(acc, fr_mul_cold(c3, d3));
  return acc;
}

// Raw sums -> evaluations at X = 0..3.  deg: 0 (raw sums are the evaluations), 2 or 3.  `claim` is
// p(0) + p(1) (ignored when have_claim is false: p(1) was summed directly and sits in raw[1]).
__device__ __noinline__ void sc_finalize(int deg, bool have_claim, const Fr* raw, const Fr& claim, Fr* out) {
  if (deg == 0) {
#pragma unroll
    for (int i = 0; i < kRaw; i++) out[i] = raw[i];
    return;
  }
  const Fr p0 = raw[0];
  const Fr p1 = have_claim ? fr_sub_cold(claim, p0) : raw[1];
  Fr p2, p3;
  if (deg == 3) {
    p2 = raw[2];
    const Fr lead = raw[3];
    Fr t = fr_add_cold(fr_add_cold(lead, lead), lead);  // 3*lead
    t = fr_add_cold(t, t);                               // 6*lead = third finite difference
    const Fr d = fr_sub_cold(p2, p1);
    p3 = fr_add_cold(fr_add_cold(t, fr_add_cold(fr_add_cold(d, d), d)), p0);  // p3 = 6 lead + 3 (p2 - p1) + p0
  } else {
    const Fr lead = raw[2];
    const Fr l2 = fr_add_cold(lead, lead);
    const Fr s = fr_sub_cold(p1, p0);                                  // c1 + c2
    const Fr p1s = fr_add_cold(p1, s);
    p2 = fr_add_cold(p1s, l2);                                         // 2 p1 - p0 + 2 c2
    p3 = fr_add_cold(fr_add_cold(p1s, s), fr_add_cold(fr_add_cold(l2, l2), l2));  // 3 p1 - 2 p0 + 6 c2
  }
  out[0] = p0;
  out[1] = p1;
  out[2] = p2;
  out[3] = p3;
}

// ---- block / grid reduction of the raw sums -----------------------------------------------------------
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

struct ScOut {
  Fr* partials;       // gridDim.x * kRaw
  unsigned* ticket;
  unsigned* next_chunk;  // streaming kernels (adjacent pairs): work counter of the 32-pair chunks handed out dynamically
  Fr* d_raw;          // kRaw raw sums of this launch (what a sharded run all-gathers)
  Fr* d_evals;        // 4 evaluations: read as "previous round" (CLAIM_PREV), then overwritten
  Fr* h_evals;        // mapped pinned host copy of d_evals
  const Fr* d_claim;  // CLAIM_HOST: p(0) + p(1) supplied by the host
  Fr* d_claim_cur;    // CLAIM_PREV: previous round polynomial at r, written by block 0 at kernel start
  int deg;            // 0 / 2 / 3
  int claim_src;
  int finalize;       // 0: sharded run (host-driven exchange), stop at the raw sums
  unsigned seq;       // finalize: written to the host mailbox (byte 128) after the evaluations, the host polls it
  unsigned long long* dbg;  // diagnostics (tuning key 7): 8 globaltimer stamps per launch at dbg + 8 * (seq & 63), or null
  PeerDev peer;       // world > 1: the raw sums of all ranks are exchanged through the NVLink mailboxes right here
};

__device__ __forceinline__ void sc_stamp(const ScOut& o, int i) {
  if (o.dbg) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    o.dbg[8 * (o.seq & 63u) + i] = t;
  }
}

__device__ __forceinline__ Fr ldcg_fr(const Fr* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  const uint4 lo4 = __ldcg(q), hi4 = __ldcg(q + 1);  // bypass L1: written by other SMs
  Fr x;
  x.v[0] = lo4.x; x.v[1] = lo4.y; x.v[2] = lo4.z; x.v[3] = lo4.w;
  x.v[4] = hi4.x; x.v[5] = hi4.y; x.v[6] = hi4.z; x.v[7] = hi4.w;
  return x;
}

// sum over the 32 lanes of a warp, valid in lane 0
__device__ __noinline__ Fr warp_sum_cold(Fr x) {
#pragma unroll 1
  for (int d = 16; d > 0; d >>= 1) x = fr_add(x, shfl_down_fr(x, d));
  return x;
}

// raw sums -> round evaluations, written to device memory and to the mapped host mailbox (one thread)
__device__ __noinline__ void sc_publish(const ScOut& o, const Fr* tot_local) {
  Fr tot[kRaw];
#pragma unroll 1
  for (int e = 0; e < kRaw; e++) tot[e] = tot_local[e];
  if (o.peer.world > 1) {
    // hypercube shards: push this rank's sums into every rank's mailbox over NVLink, wait for everybody's,
    // fold the rows in rank order -- every rank ends the launch holding the same global round polynomial
    peer_publish(o.peer, tot_local, kRaw);
    if (!peer_wait(o.peer)) return;
#pragma unroll 1
    for (int e = 0; e < kRaw; e++) {
      Fr acc = Fr::zero();
#pragma unroll 1
      for (int g = 0; g < o.peer.world; g++) acc = fr_add_cold(acc, peer_load(o.peer, g, e));
      tot[e] = acc;
    }
  }
#pragma unroll 1
  for (int e = 0; e < kRaw; e++) st_fr(&o.d_raw[e], tot[e]);
  if (o.finalize) {
    Fr claim = Fr::zero();
    if (o.claim_src == CLAIM_PREV) claim = ldcg_fr(o.d_claim_cur);
    else if (o.claim_src == CLAIM_HOST) claim = ld_fr(o.d_claim);
    Fr ev[4];
    sc_finalize(o.deg, o.claim_src != CLAIM_NONE, tot, claim, ev);
    sc_stamp(o, 4);
#pragma unroll 1
    for (int e = 0; e < 4; e++) {
      st_fr(&o.d_evals[e], ev[e]);
      st_fr(&o.h_evals[e], ev[e]);
    }
    // the host does not wait for the kernel to retire: it spins on this word (zolt_sc_round_evals)
    __threadfence_system();
    *reinterpret_cast<volatile unsigned*>(o.h_evals + 4) = o.seq;
  }
  __threadfence_system();
  sc_stamp(o, 5);
}

// Each block publishes its partial; the last block to arrive folds all partials, turns the raw sums into
// the round evaluations and writes them to device memory and to the mapped host mailbox.
template <int NE>
__device__ __forceinline__ void grid_finish(Fr* acc, Fr* smem, const ScOut& o, const Fr& r) {
  __shared__ bool is_last;
  block_reduce<NE>(acc, smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int e = 0; e < NE; e++) st_fr(&o.partials[(size_t)blockIdx.x * NE + e], acc[e]);
    __threadfence();
    unsigned t = atomicAdd(o.ticket, 1u);
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
    for (int e = 0; e < NE; e++) tot[e] = fr_add(tot[e], ldcg_fr(&o.partials[(size_t)b * NE + e]));
  }
  block_reduce<NE>(tot, smem);
  if (threadIdx.x == 0) {
    Fr raw4[kRaw];
#pragma unroll
    for (int e = 0; e < kRaw; e++) raw4[e] = e < NE ? tot[e < NE ? e : 0] : Fr::zero();
    *o.ticket = 0;
    sc_publish(o, raw4);
  }
}

// Grid-level fold of the raw sums and publication.  `x` is the block's sum number `warp`, valid in lane 0 of
// warps 0..kRaw-1 (a block may have more warps; they only take part in the barriers).  Each block stores its
// four sums and takes a ticket; the last block to arrive folds all partials (warp e owns sum e) and publishes.
__device__ __noinline__ void grid_fold_publish(Fr x, const ScOut& o) {
  __shared__ bool is_last;
  __shared__ Fr fin[kRaw];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (gridDim.x > 1) {
    if (warp < kRaw && lane == 0) {
      st_fr(&o.partials[(size_t)blockIdx.x * kRaw + warp], x);
      __threadfence();
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      const unsigned t = atomicAdd(o.ticket, 1u);
      is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (tid == 0) sc_stamp(o, 2);
    if (warp < kRaw) {
      Fr tot = Fr::zero();
#pragma unroll 1
      for (unsigned b = lane; b < gridDim.x; b += 32) tot = fr_add(tot, ldcg_fr(&o.partials[(size_t)b * kRaw + warp]));
      x = warp_sum_cold(tot);
    }
  }
  if (warp < kRaw && lane == 0) fin[warp] = x;
  __syncthreads();
  if (tid == 0) {
    if (gridDim.x > 1) *o.ticket = 0;
    if (o.next_chunk) *o.next_chunk = 0;  // every block took its last chunk before it took its ticket
    sc_stamp(o, 3);
    sc_publish(o, fin);
  }
}

// The streaming round kernels' epilogue (kRoundThreads = 4 warps, kRaw = 4 sums): warp e owns sum e.  `vals` is the
// block's dynamic scratch, re-used as [kRaw][kRoundThreads] field elements once every thread has left its loop.
__device__ __noinline__ void round_finish(const Fr* raw, Fr* vals, const ScOut& o) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (blockIdx.x == 0 && tid == 0) sc_stamp(o, 1);
  if (o.dbg && tid == 0) {  // per block: when it left its loop, and on which SM (words 512 + 2 b, overwritten by every launch)
    unsigned long long tt;
    unsigned sm;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    o.dbg[512 + 2 * blockIdx.x] = tt;
    o.dbg[512 + 2 * blockIdx.x + 1] = sm;
  }
  __syncthreads();
#pragma unroll 1
  for (int e = 0; e < kRaw; e++) st_fr(&vals[e * kRoundThreads + tid], raw[e]);
  __syncthreads();
  Fr x = vals[warp * kRoundThreads + lane];
#pragma unroll 1
  for (int j = 1; j < kRoundThreads / 32; j++) x = fr_add(x, vals[warp * kRoundThreads + lane + 32 * j]);
  x = warp_sum_cold(x);
  grid_fold_publish(x, o);
}

// ---- the round kernel ------------------------------------------------------------------------------
// MODE 0: evaluate the round polynomial of the current tables (src, length n), p(1) summed directly.
// MODE 1: the same, p(1) from the claim the host supplied (zolt_sc_set_claim).
// MODE 2: FUSED -- bind src (length n) to r into dst (length n/2) and evaluate the NEXT round's
//         polynomial over dst in the same pass; its claim is the previous polynomial at r.  n_extra tables
//         after the K combiner tables are bound only.
//
// BIND selects the multiplier pipe of the bind (MODE 2 only): 0 = integer (fp_bind_fixed, wide IMADs), 1 = FP64
// (fp_bind_fixed_f64: 128 DFMAs per bound element on the otherwise idle FP64 pipe; the combiner's products keep
// the integer multiplier to themselves).  Both produce the same canonical elements.
struct BindTables {
  FixedMulTable i32;
  // Two identical copies, used alternately by successive binds: the table (128 doubles) does not fit the 63 uniform
  // registers, and a loop-invariant table makes ptxas hoist all of it out of the bind loop and spill it to local memory
  // (r2 v1: LDL + R2UR per DFMA operand, the pass ran 1.6x SLOWER than the integer bind).  An address that changes with
  // the loop counter keeps the LDCU loads next to their DFMAs.
  FixedMulTableF64 f64[2];
  unsigned f64_mask;  // bit k: table k's binds take the FP64 pipe (BIND == 1 kernels)
};

// what a launch carries: the integer kernels only the 256-byte integer table (2 KB less to marshal per launch)
struct BindTableI32 {
  FixedMulTable i32;
};
template <int BIND> struct BindArg { using type = BindTableI32; };
template <> struct BindArg<1> { using type = BindTables; };

template <int COMB, int PAIR, int MODE, int BIND = 0>
__global__ void __launch_bounds__(kRoundThreads, MODE == 2 ? kFusedBlocksPerSm : kEvalBlocksPerSm)
sc_round_kernel(const __grid_constant__ ScPtrs t, int n_extra, size_t n, const __grid_constant__ Fr r, SplitEqDev sq,
                const __grid_constant__ ScOut out, const __grid_constant__ typename BindArg<BIND>::type rt,
                const __grid_constant__ TmapArray tm) {
  constexpr int K = CombTraits<COMB>::K;
  constexpr int NW = CombTraits<COMB>::NW;
  constexpr bool FUSED = MODE == 2;
  constexpr bool P1 = MODE == 0 && CombTraits<COMB>::DEG != 0;
  constexpr int NACC = NW + (P1 ? 1 : 0);
  extern __shared__ __align__(128) uint4 scratch[];  // dynamic: exceeds the 48 KB static limit
  if constexpr (FUSED) {
    // this round's claim p(0) + p(1) is the previous round polynomial at r: one thread evaluates it while the
    // grid streams the tables, so it is off the serial tail of the launch
    if (blockIdx.x == 0 && threadIdx.x == 0 && out.finalize) {
      Fr prev[4];
#pragma unroll 1
      for (int e = 0; e < 4; e++) prev[e] = ld_fr(&out.d_evals[e]);
      st_fr(out.d_claim_cur, newton_eval4(prev, r));
      __threadfence();
    }
  }
  if (out.dbg && blockIdx.x == 0 && threadIdx.x == 0) sc_stamp(out, 0);
  uint32_t wacc[NACC > 0 ? NACC : 1][16];
  Fr facc[2] = {Fr::zero(), Fr::zero()};
#pragma unroll
  for (int i = 0; i < NACC; i++) wide_zero(wacc[i]);

  const size_t n_cur = FUSED ? n / 2 : n;  // length of the table the evaluations are taken over
  const size_t pairs = n_cur / 2;
  const size_t half_old = n / 2;           // HALF pairing: distance between bind partners in src
  const size_t half_cur = n_cur / 2;       // HALF pairing: distance between eval partners
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  size_t g0 = blockIdx.x * (size_t)blockDim.x + (threadIdx.x & ~31);  // first pair of this warp's 32
  // Half-split pairing: the trip count is the same for every thread of the grid (a ragged last iteration is predicated), so
  // the loops are warp-convergent code: warp-cooperative copies and __syncwarp are safe.  (The adjacent-pair loops below take
  // their chunks from a counter instead; a warp's lanes still share every chunk index.)
  const size_t iters = (pairs + stride - 1) / stride;

  if constexpr (PAIR == ZOLT_PAIR_ADJACENT) {
    // ============ adjacent pairs (LowToHigh): TMA boxes of 32 rows per warp and table ===================================
    // 1024-byte aligned start (the swizzle pattern is a function of the shared-memory address); plain pointer arithmetic on
    // `scratch` so that the compiler keeps the shared address space (LDS/STS, not generic LD/ST)
    unsigned char* const smem = reinterpret_cast<unsigned char*>(scratch) + ((1024u - (smem_u32(scratch) & 1023u)) & 1023u);
    // Work is handed out in chunks of 32 pairs (one TMA box row range per warp) from a grid-wide counter, not by a fixed
    // stride: the SM's warp arbiter is not fair (r2 v14 per-block stamps: of the 3-4 co-resident blocks of an SM two finished
    // a statically assigned share at 60-70 % of the kernel's duration and the starved one ran the rest alone, one warp per
    // scheduler, with the multiplier half idle).  Every warp now stays busy until the table runs out.  A warp's first chunk
    // is static; the next two are always known one iteration ahead so the copy of chunk i+1 overlaps the arithmetic of chunk i
    // and the counter's round trip is never waited for.  (Field sums are exact: the order of the terms does not matter.)
    const unsigned nchunks = (unsigned)((pairs + 31) / 32);
    const unsigned grid_warps = gridDim.x * (kRoundThreads / 32);
    auto grab = [&]() -> unsigned {  // valid in lane 0; broadcast where it is consumed
      unsigned v = 0;
      if (lane == 0) v = grid_warps + atomicAdd(out.next_chunk, 1u);
      return v;
    };
    unsigned cur = blockIdx.x * (kRoundThreads / 32) + warp;
    unsigned nxt = __shfl_sync(0xffffffffu, grab(), 0);
    if constexpr (FUSED) {
      // per warp: one 4 KB slot for a table's sources (32 rows x 128 B, 128-byte swizzle), then the stash of the bound
      // values ([unit][lane] 16-byte slots, unit 4k = lo, 4k+2 = hi of table k); mbarriers after the last warp
      constexpr int kSlot = 32 * 128, kWarpBytes = kSlot + 4 * K * 32 * 16;
      unsigned char* const slot = smem + warp * kWarpBytes;
      uint4* const stash = reinterpret_cast<uint4*>(slot + kSlot) + lane;
      const unsigned char* const row = slot + lane * 128;
      const int sw = (lane & 7) << 4;  // chunk c of this lane's row sits at byte (16 c) ^ sw
      unsigned long long* const bar = reinterpret_cast<unsigned long long*>(smem + (kRoundThreads / 32) * kWarpBytes) + warp;
      if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      }
      __syncwarp();
      const int kt = K + n_extra;
      // `dep` is always zero but computed from the words just read out of the slot: the copy that overwrites the slot cannot
      // issue before those shared loads have returned (the TMA unit is not ordered behind the LSU queue the way cp.async is)
      auto issue = [&](unsigned chunk, int k, unsigned dep) {
        if (lane == 0 && chunk < nchunks) {
          mbar_expect_tx(bar, kSlot);
          tma_rows_g2s(slot, &tm.m[k], (int)(chunk * 32u + dep), bar);  // row g = sources 4g .. 4g+3 of pair g
        }
      };
      auto ld_chunks = [&](int c) {  // the element made of chunks c, c+1 of this lane's row
        const uint4 a = *reinterpret_cast<const uint4*>(row + ((16 * c) ^ sw));
        const uint4 b = *reinterpret_cast<const uint4*>(row + ((16 * (c + 1)) ^ sw));
        Fr x;
        x.v[0] = a.x; x.v[1] = a.y; x.v[2] = a.z; x.v[3] = a.w;
        x.v[4] = b.x; x.v[5] = b.y; x.v[6] = b.z; x.v[7] = b.w;
        return x;
      };
      issue(cur, 0, 0u);
      unsigned parity = 0;
      while (cur < nchunks) {
        const size_t g = (size_t)cur * 32 + lane;
        const bool act = g < pairs;
        const unsigned after = grab();  // in flight while this chunk is processed
#pragma unroll 1
        for (int k = 0; k < kt; k++) {
          mbar_wait(bar, parity);  // this table's sources have landed
          parity ^= 1u;
          const Fr a0 = ld_chunks(0), a1 = ld_chunks(2), b0 = ld_chunks(4), b1 = ld_chunks(6);  // zeros when !act
          __syncwarp();  // every lane has read its row before the slot is re-armed
          {
            // canonical limbs: the top word of every element is below 2^30, so this is 0
            const unsigned dep = (a0.v[7] | a1.v[7] | b0.v[7] | b1.v[7] | (a0.v[3] & a1.v[3] & b0.v[3] & b1.v[3] & a0.v[7])) >> 31;
            const bool wrap = k + 1 == kt;
            issue(wrap ? nxt : cur, wrap ? 0 : k + 1, dep);
          }
          Fr v0, v1;  // lo + r (hi - lo): multiplier table in the constant bank
          if constexpr (BIND == 1) {
            // bit k of the mask: this table's binds run on the FP64 pipe, else on the integer multiplier
            if ((rt.f64_mask >> k) & 1u) {
              fp_bind2_fixed_f64<FrParams>(a0, a1, b0, b1, rt.f64[k & 1], v0, v1);
            } else {
              v0 = fp_bind_fixed<FrParams>(a0, a1, rt.i32);
              v1 = fp_bind_fixed<FrParams>(b0, b1, rt.i32);
            }
          } else {
            v0 = fp_bind_fixed<FrParams>(a0, a1, rt.i32);
            v1 = fp_bind_fixed<FrParams>(b0, b1, rt.i32);
          }
          if (act) {
            st_fr(t.dst[k] + 2 * g, v0);
            st_fr(t.dst[k] + 2 * g + 1, v1);
          }
          if (k < K) {
            stash[(4 * k) * 32] = make_uint4(v0.v[0], v0.v[1], v0.v[2], v0.v[3]);
            stash[(4 * k + 1) * 32] = make_uint4(v0.v[4], v0.v[5], v0.v[6], v0.v[7]);
            stash[(4 * k + 2) * 32] = make_uint4(v1.v[0], v1.v[1], v1.v[2], v1.v[3]);
            stash[(4 * k + 3) * 32] = make_uint4(v1.v[4], v1.v[5], v1.v[6], v1.v[7]);
          }
        }
        if (act) {
          Fr w;
          if constexpr (CombTraits<COMB>::GRUEN) w = sq.weight(g);
          accumulate_pair<COMB, P1>(wacc, facc, UnitStash<32>{stash}, w);
        }
        cur = nxt;
        nxt = __shfl_sync(0xffffffffu, after, 0);
      }
    } else {
      // evaluation only: per warp two buffers of K boxes (32 rows x 64 B, 64-byte swizzle), filled one iteration ahead
      constexpr int kTab = 32 * 64, kBuf = K * kTab, kWarpBytes = 2 * kBuf;
      unsigned char* const wbase = smem + warp * kWarpBytes;
      unsigned long long* const bars = reinterpret_cast<unsigned long long*>(smem + (kRoundThreads / 32) * kWarpBytes) + 2 * warp;
      if (lane == 0) {
        mbar_init(bars, 1);
        mbar_init(bars + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      }
      __syncwarp();
      auto issue = [&](unsigned chunk, int buf) {
        if (lane == 0 && chunk < nchunks) {
          mbar_expect_tx(bars + buf, kBuf);
#pragma unroll
          for (int k = 0; k < K; k++) tma_rows_g2s(wbase + buf * kBuf + k * kTab, &tm.m[k], (int)(chunk * 32u), bars + buf);  // row g = (lo, hi) of pair g
        }
      };
      issue(cur, 0);
      unsigned parity[2] = {0u, 0u};
      int buf = 0;
      const int sw = ((lane >> 1) & 3) << 4;  // 64-byte swizzle: chunk c of row `lane` sits at byte (16 c) ^ sw
      while (cur < nchunks) {
        const size_t g = (size_t)cur * 32 + lane;
        const unsigned after = grab();  // in flight while this chunk is processed
        issue(nxt, buf ^ 1);
        mbar_wait(bars + buf, parity[buf]);
        parity[buf] ^= 1u;
        if (g < pairs) {
          Fr w;
          if constexpr (CombTraits<COMB>::GRUEN) w = sq.weight(g);
          const SwizzledRows64 st{wbase + buf * kBuf + lane * 64, sw, kTab};
          accumulate_pair<COMB, P1>(wacc, facc, st, w);
        }
        __syncwarp();  // every lane is done with this buffer before the next iteration re-arms it
        buf ^= 1;
        cur = nxt;
        nxt = __shfl_sync(0xffffffffu, after, 0);
      }
    }
  } else {
    // ============ half-split pairs (HighToLow): cp.async into [unit][thread] slots =====================================
    // (degree-1 provers; their sources are four / two separate runs per table, copied cooperatively 512 contiguous bytes per
    // warp instruction, lane L the 16-byte piece L, each piece dropped into the slot of the thread that owns it)
    uint4* const scr = scratch + threadIdx.x;
    uint4* const warp_scr = scratch + (threadIdx.x & ~31);  // slot of lane 0 of this warp
    size_t g = g0 + lane;
    if constexpr (FUSED) {
      // units 0..15: stash of the bound values (unit 4k lo, 4k+2 hi of table k); units 16..23: the four source elements
      // of one table's two binds
      const int kt = K + n_extra;
      auto prefetch_table = [&](size_t gi, int k) {
        if (gi >= pairs) return;
        const Fr* s = t.src[k];
        const int vp = pairs - gi < 32 ? (int)(pairs - gi) : 32;  // valid pairs of this warp's chunk
        // outputs g, g + half_cur <- sources (g, g + half_old), (g + half_cur, g + half_cur + half_old): four runs
        warp_copy_run<2>(warp_scr, 16, s + gi, 2 * vp, lane);
        warp_copy_run<2>(warp_scr, 18, s + gi + half_old, 2 * vp, lane);
        warp_copy_run<2>(warp_scr, 20, s + gi + half_cur, 2 * vp, lane);
        warp_copy_run<2>(warp_scr, 22, s + gi + half_cur + half_old, 2 * vp, lane);
      };
      prefetch_table(g0, 0);
      cp_async_commit();
      for (size_t it = 0; it < iters; it++, g += stride, g0 += stride) {
        const bool act = g < pairs;
#pragma unroll 1
        for (int k = 0; k < kt; k++) {
          cp_async_wait<0>();  // this table's sources have landed (every lane waits for the pieces IT copied) ...
          __syncwarp();        // ... and the pieces the other lanes copied into this thread's slots are visible
          const Fr a0 = scr_ld(scr, 16), a1 = scr_ld(scr, 18), b0 = scr_ld(scr, 20), b1 = scr_ld(scr, 22);
          __syncwarp();        // every lane has read its slots before any lane re-arms them
          {
            const bool wrap = k + 1 == kt;
            prefetch_table(wrap ? g0 + stride : g0, wrap ? 0 : k + 1);
            cp_async_commit();
          }
          const Fr v0 = fp_bind_fixed<FrParams>(a0, a1, rt.i32), v1 = fp_bind_fixed<FrParams>(b0, b1, rt.i32);
          if (act) {
            st_fr(t.dst[k] + g, v0);
            st_fr(t.dst[k] + g + half_cur, v1);
          }
          if (k < K) {
            scr_st(scr, 4 * k, v0);
            scr_st(scr, 4 * k + 2, v1);
          }
        }
        if (act) accumulate_pair<COMB, P1>(wacc, facc, UnitStash<kScrStride>{scr}, Fr::zero());
      }
    } else {
      auto prefetch_pairs = [&](size_t gi, int buf) {
        if (gi >= pairs) return;
        const int vp = pairs - gi < 32 ? (int)(pairs - gi) : 32;
#pragma unroll
        for (int k = 0; k < K; k++) {
          warp_copy_run<2>(warp_scr, 16 * buf + 4 * k, t.src[k] + gi, 2 * vp, lane);
          warp_copy_run<2>(warp_scr, 16 * buf + 4 * k + 2, t.src[k] + gi + half_cur, 2 * vp, lane);
        }
      };
      prefetch_pairs(g0, 0);
      cp_async_commit();
      int buf = 0;
      for (size_t it = 0; it < iters; it++, g += stride, g0 += stride) {
        prefetch_pairs(g0 + stride, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();  // the other lanes' pieces of this buffer are visible
        if (g < pairs) accumulate_pair<COMB, P1>(wacc, facc, UnitStash<kScrStride>{scr + 16 * buf * kScrStride}, Fr::zero());
        __syncwarp();  // every lane is done with this buffer before the next iteration re-arms it
        buf ^= 1;
      }
    }
    cp_async_wait<0>();
  }
  Fr raw[kRaw];
  collapse<COMB, P1>(wacc, facc, raw);
  round_finish(raw, reinterpret_cast<Fr*>(scratch), out);
}

// ---- small rounds --------------------------------------------------------------------------------------
// Below ~2^14 pairs the streaming kernel is pure latency: one thread walks 2k binds, the whole combiner and
// three reductions in sequence (~25 us whatever the size; r1 v4 launch list).  This kernel spreads the same
// arithmetic over threads instead: ONE bound element per thread, then ONE (pair, sample point) term per
// thread, every field primitive called out of line so the code stays a few KB.  The raw sums it produces are
// the same field elements (sums of the same canonical terms), so it feeds sc_publish unchanged.
constexpr int kSmallWorkWarps = 4;                            // warp e evaluates raw sum e
constexpr int kSmallThreads = 32 * (kSmallWorkWarps + 1);     // + one warp for the claim p_prev(r)
constexpr int kSmallPairs = 16;                               // pairs of the current table per chunk
constexpr int kSmallTables = 5;                               // most combiner tables of any combiner (Stage-4 dense)

struct SmallArgs {
  int comb, k, n_extra, deg, fused, p1_direct;
};

__device__ __noinline__ Fr bind_pair_cold(Fr lo, Fr hi, Fr r) { return fr_add(lo, fr_mul(r, fr_sub(hi, lo))); }

enum { PT_NONE = 0, PT_0, PT_1, PT_2, PT_LEAD, PT_3 };

// the term of pair p of this chunk that belongs to raw sum e (layout of the raw sums: see the combiners above)
template <int NP>
__device__ __noinline__ Fr small_term(const SmallArgs& a, int e, const Fr (*stash)[2][NP], int p, const SplitEqDev& sq,
                                      size_t g) {
  const bool gruen = a.comb == ZOLT_SC_GRUEN_AB || a.comb == ZOLT_SC_GRUEN_AB_MINUS_C || a.comb == ZOLT_SC_STAGE4_GRUEN;
  int pt = PT_NONE;
  if (a.deg == 3) pt = e == 0 ? PT_0 : (e == 1 ? (a.p1_direct ? PT_1 : PT_NONE) : (e == 2 ? PT_2 : PT_LEAD));
  else if (a.deg == 2) pt = e == 0 ? PT_0 : (e == 1 ? (a.p1_direct ? PT_1 : PT_NONE) : (e == 2 ? PT_LEAD : PT_NONE));
  else if (gruen) pt = e == 0 ? PT_0 : (e == 1 ? PT_LEAD : PT_NONE);
  else if (a.comb == ZOLT_SC_STAGE4_DENSE || a.comb == ZOLT_SC_EQ_DOT2) pt = e == 0 ? PT_0 : (e == 1 ? PT_2 : (e == 2 ? PT_3 : PT_NONE));
  else pt = e == 0 ? PT_0 : (e == 1 ? PT_1 : PT_NONE);
  if (pt == PT_NONE) return Fr::zero();
  Fr v[kSmallTables];
#pragma unroll 1
  for (int k = 0; k < a.k; k++) {
    const Fr lo = stash[k][0][p], hi = stash[k][1][p];
    if (pt == PT_0) v[k] = lo;
    else if (pt == PT_1) v[k] = hi;
    else {
      const Fr d = fr_sub_cold(hi, lo);
      v[k] = pt == PT_LEAD ? d : fr_add_cold(hi, d);
      if (pt == PT_3) v[k] = fr_add_cold(v[k], d);
    }
  }
  const bool lead = pt == PT_LEAD;
  switch (a.comb) {
    case ZOLT_SC_SUM: return v[0];
    case ZOLT_SC_PROD2: return fr_mul_cold(v[0], v[1]);
    case ZOLT_SC_PROD3: return fr_mul_cold(fr_mul_cold(v[0], v[1]), v[2]);
    case ZOLT_SC_EQ_A_BMC: return fr_mul_cold(fr_mul_cold(v[0], v[1]), fr_sub_cold(v[2], v[3]));
    case ZOLT_SC_EQ_AB_MINUS_C: {
      Fr t = fr_mul_cold(v[1], v[2]);
      if (!lead) t = fr_sub_cold(t, v[3]);  // c(X) is linear: no cubic term
      return fr_mul_cold(v[0], t);
    }
    case ZOLT_SC_STAGE4_DENSE:  // eq * (ra val + wa (val + inc)) at the sample point
      return fr_mul_cold(v[0], fr_add_cold(fr_mul_cold(v[1], v[3]), fr_mul_cold(v[2], fr_add_cold(v[3], v[4]))));
    case ZOLT_SC_EQ_DOT2:  // w * (a b + c d) at the sample point
      return fr_mul_cold(v[0], fr_add_cold(fr_mul_cold(v[1], v[2]), fr_mul_cold(v[3], v[4])));
    case ZOLT_SC_STAGE4_GRUEN: {  // weight * (ra val + wa (val + inc))
      const size_t x_out = (g >> sq.in_bits) & sq.out_mask, x_in = g & ((size_t(1) << sq.in_bits) - 1);
      const Fr w = fr_mul_cold(ld_fr(sq.e_out + x_out), ld_fr(sq.e_in + x_in));
      return fr_mul_cold(w, fr_add_cold(fr_mul_cold(v[0], v[2]), fr_mul_cold(v[1], fr_add_cold(v[2], v[3]))));
    }
    default: {  // Gruen: weight * (l r [- c])
      const size_t x_out = (g >> sq.in_bits) & sq.out_mask, x_in = g & ((size_t(1) << sq.in_bits) - 1);
      const Fr w = fr_mul_cold(ld_fr(sq.e_out + x_out), ld_fr(sq.e_in + x_in));
      Fr t = fr_mul_cold(v[0], v[1]);
      if (a.comb == ZOLT_SC_GRUEN_AB_MINUS_C && !lead) t = fr_sub_cold(t, v[2]);
      return fr_mul_cold(w, t);
    }
  }
}

template <int PAIR>
__global__ void __launch_bounds__(kSmallThreads)
sc_small_kernel(const __grid_constant__ ScPtrs t, const SmallArgs a, size_t n, const __grid_constant__ Fr r, const SplitEqDev sq,
                const __grid_constant__ ScOut out) {
  __shared__ Fr stash[kSmallTables][2][kSmallPairs];  // [table][half][pair of the chunk]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (a.fused && warp == kSmallWorkWarps && blockIdx.x == 0 && lane == 0 && out.finalize) {
    Fr prev[4];
#pragma unroll 1
    for (int e = 0; e < 4; e++) prev[e] = ld_fr(&out.d_evals[e]);
    st_fr(out.d_claim_cur, newton_eval4(prev, r));
    __threadfence();
  }
  const size_t n_cur = a.fused ? n / 2 : n;
  const size_t pairs = n_cur / 2, half_old = n / 2, half_cur = n_cur / 2;
  const int kt = a.k + a.n_extra;
  const size_t n_chunks = (pairs + kSmallPairs - 1) / kSmallPairs;
  Fr x = Fr::zero();
  for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    if (warp < kSmallWorkWarps) {
#pragma unroll 1
      for (int it = tid; it < kSmallPairs * 2 * kt; it += 32 * kSmallWorkWarps) {
        const int p = it % kSmallPairs, h = (it / kSmallPairs) & 1, k = it / (2 * kSmallPairs);
        const size_t g = c * kSmallPairs + p;
        if (g >= pairs) continue;
        const size_t idx = (PAIR == ZOLT_PAIR_ADJACENT) ? 2 * g + h : g + (h ? half_cur : 0);
        Fr v;
        if (a.fused) {
          const Fr* s = t.src[k];
          const Fr lo = (PAIR == ZOLT_PAIR_ADJACENT) ? ld_fr(s + 2 * idx) : ld_fr(s + idx);
          const Fr hi = (PAIR == ZOLT_PAIR_ADJACENT) ? ld_fr(s + 2 * idx + 1) : ld_fr(s + idx + half_old);
          v = bind_pair_cold(lo, hi, r);
          st_fr(t.dst[k] + idx, v);
        } else {
          v = ld_fr(t.src[k] + idx);
        }
        if (k < a.k) stash[k][h][p] = v;
      }
    }
    __syncthreads();
    if (warp < kSmallWorkWarps && lane < kSmallPairs) {
      const size_t g = c * kSmallPairs + lane;
      if (g < pairs) x = fr_add_cold(x, small_term(a, warp, stash, lane, sq, g));
    }
    __syncthreads();
  }
  if (warp < kSmallWorkWarps) x = warp_sum_cold(x);
  grid_fold_publish(x, out);
}

// ---- persistent tail ------------------------------------------------------------------------------------------------
// Below ~2^12 pairs a round is all fixed cost: ~25 us of kernel (launch, one pair per thread, grid fold, finalize) plus
// ~13 us of host (Keccak, launch, mailbox round trip), 12 times per sumcheck (r2 v9 launch list) -- more than the
// streaming rounds 3..23 together.  The tail therefore runs as ONE kernel that stays resident for all remaining
// rounds: a thread-block CLUSTER (8 CTAs, co-scheduled by the hardware, so its barrier cannot dead-lock on residency the
// way a hand-made grid barrier can) that, per round,
//   waits for the host's command (the challenge) in mapped pinned memory -> binds every table -> evaluates the next round
//   polynomial -> cluster barrier -> CTA 0 folds the 8 partials, completes the polynomial and writes it, then the
//   command's sequence number, into the mapped result mailbox the host is already spinning on.
// The call pattern of the reference (SumcheckInstance.computeRoundPolyFn / bindChallengeFn, prover.zig:405-428) is
// unchanged: zolt_sc_bind posts the command, zolt_sc_round_evals reads the mailbox.  The kernel runs on a private
// stream of the instance, so other work on the context's stream (another instance of a batched sumcheck, an MSM) is
// not queued behind it; anything that needs the tables (download, table_dev, free) first tells it to exit.
constexpr int kTailCluster = 8;
constexpr int kTailThreads = 512;
constexpr int kTailChunk = 128;  // pairs per CTA step (stash: 4 tables x 2 halves x 128 x 32 B = 32 KB)
enum { TAIL_OP_BIND = 1, TAIL_OP_EXIT = 2 };
enum { TAIL_ST_RUNNING = 1, TAIL_ST_DONE = 2, TAIL_ST_TIMEOUT = 3 };
// layout of the instance's 256-byte mapped mailbox: [0,128) evaluations, 128 result sequence number, 132 tail status,
// 136 number of the last command the kernel has taken (the single command slot may be rewritten after that) (device -> host); 160 command sequence number, 164 op, 168 result sequence number to report, [192, 224) challenge
// (host -> device)
constexpr int kMbResultSeq = 128, kMbTailStatus = 132, kMbCmdAck = 136, kMbCmd = 144, kCmdChunks = 6;  // command: 6 chunks of 16 bytes

struct TailArgs {
  const Fr* in[kMaxTables];  // tables of the instance: inputs, and the two ping-pong sets
  Fr* b0[kMaxTables];
  Fr* b1[kMaxTables];
  int cur;                   // -1 / 0 / 1 as in zolt_sc
  size_t n;                  // current table length when the kernel starts
  SmallArgs a;
  const Fr* e_out;           // split-eq prefix tables (Gruen combiners), level L at offset 2^L - 1
  const Fr* e_in;
  int n_tau;
  int eq_skip;               // high bits of the table index the split-eq weights do not depend on (Stage 4: the row k)
  Fr* partials;              // kTailCluster * kRaw
  Fr* d_evals;               // previous round's evaluations in, latest evaluations out
  Fr* d_raw;
  unsigned char* mailbox;    // device address of the mapped mailbox
  unsigned long long timeout_ns;        // patience for every command but the first
  unsigned long long first_timeout_ns;  // ... and for the first, which the launching call posts at once
  unsigned long long* dbg;   // optional: per round 4 globaltimer stamps of CTA 0 (command seen, work done, folded, published)
};

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

// sc_finalize with the claim, arithmetic inlined (the tail kernel's one serial step per round)
__device__ __forceinline__ void sc_finalize_inl(int deg, const Fr* raw, const Fr& claim, Fr* out) {
  if (deg == 0) {
#pragma unroll
    for (int i = 0; i < kRaw; i++) out[i] = raw[i];
    return;
  }
  const Fr p0 = raw[0];
  const Fr p1 = fr_sub(claim, p0);
  Fr p2, p3;
  if (deg == 3) {
    p2 = raw[2];
    const Fr lead = raw[3];
    Fr t = fr_add(fr_add(lead, lead), lead);
    t = fr_add(t, t);
    const Fr d = fr_sub(p2, p1);
    p3 = fr_add(fr_add(t, fr_add(fr_add(d, d), d)), p0);
  } else {
    const Fr lead = raw[2];
    const Fr l2 = fr_add(lead, lead);
    const Fr s = fr_sub(p1, p0);
    const Fr p1s = fr_add(p1, s);
    p2 = fr_add(p1s, l2);
    p3 = fr_add(fr_add(p1s, s), fr_add(fr_add(l2, l2), l2));
  }
  out[0] = p0;
  out[1] = p1;
  out[2] = p2;
  out[3] = p3;
}

// store a 32-bit word into the same shared-memory variable of CTA `rank` of this cluster (distributed shared memory)
__device__ __forceinline__ void dsmem_st_u32(const void* local_var, unsigned rank, unsigned v) {
  unsigned remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_var)), "r"(rank));
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote), "r"(v) : "memory");
}

template <int PAIR>
__global__ void __cluster_dims__(kTailCluster, 1, 1) __launch_bounds__(kTailThreads)
sc_tail_kernel(const __grid_constant__ TailArgs ta) {
  __shared__ Fr stash[4][2][kTailChunk];
  __shared__ Fr red[kTailThreads / 32];
  __shared__ unsigned cmd_sh[2 + 8];  // op, result seq, challenge: written by CTA 0 into every CTA's copy
  __shared__ unsigned claim_w[8];     // CTA 0: the host-supplied claim of the round being evaluated
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned cta = cluster_ctarank();
  const SmallArgs& a = ta.a;
  const int kt = a.k + a.n_extra;
  int cur = ta.cur;
  size_t n = ta.n;
  unsigned expect = 1;
  volatile unsigned* const mb32 = reinterpret_cast<volatile unsigned*>(ta.mailbox);
  unsigned status = TAIL_ST_DONE;
  // every CTA of the cluster has started before anybody stores into a peer's shared memory (the hardware co-schedules a
  // cluster, but only this barrier makes the peers' shared memory valid to address: compute-sanitizer flags the store otherwise)
  cluster_sync_all();
  while (n >= 2) {
    // ---- the host's next command: ONE thread of the cluster polls the mapped word (a PCIe read per poll) and hands the
    // command to every CTA through distributed shared memory; everybody else waits at the cluster barrier ----
    if (cta == 0 && tid == 0) {
      unsigned long long t0;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      unsigned w[18];
      const volatile uint4* const cmd = reinterpret_cast<const volatile uint4*>(ta.mailbox + kMbCmd);
      for (unsigned spins = 1;; spins++) {
        uint4 c[kCmdChunks];  // independent system-memory loads, in flight together
#pragma unroll
        for (int i = 0; i < kCmdChunks; i++)
          asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(c[i].x), "=r"(c[i].y), "=r"(c[i].z), "=r"(c[i].w) : "l"(cmd + i));
        bool ok = true;
#pragma unroll
        for (int i = 0; i < kCmdChunks; i++) ok = ok && c[i].x == expect;
        if (ok) {
#pragma unroll
          for (int i = 0; i < kCmdChunks; i++) {
            w[3 * i] = c[i].y;
            w[3 * i + 1] = c[i].z;
            w[3 * i + 2] = c[i].w;
          }
          break;
        }
        if ((spins & 63u) == 0) {
          unsigned long long t1;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
          if (t1 - t0 > (expect == 1 ? ta.first_timeout_ns : ta.timeout_ns)) {
            w[0] = 0xffffffffu;
#pragma unroll
            for (int i = 1; i < 18; i++) w[i] = 0;
            break;
          }
        }
      }
      // CTA 0 alone needs the claim; every CTA needs op, result seq and the challenge
#pragma unroll
      for (int i = 0; i < 8; i++) claim_w[i] = w[10 + i];
#pragma unroll 1
      for (unsigned b = 0; b < kTailCluster; b++) {
#pragma unroll
        for (int i = 0; i < 10; i++) dsmem_st_u32(&cmd_sh[i], b, w[i]);
      }
    }
    __syncwarp();
    cluster_sync_all();  // the command sits in every CTA's shared memory
    const unsigned op = cmd_sh[0];
    if (cta == 0 && tid == 0 && op != 0xffffffffu) mb32[kMbCmdAck / 4] = expect;  // taken: the host may rewrite the command slot (a time-out acknowledges nothing)
    if (ta.dbg && cta == 0 && tid == 0) { unsigned long long tt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt)); ta.dbg[(expect - 1) * 4 + 0] = tt; }
    if (op != TAIL_OP_BIND) {  // uniform over the cluster
      status = op == TAIL_OP_EXIT ? TAIL_ST_DONE : TAIL_ST_TIMEOUT;
      break;
    }
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = cmd_sh[2 + i];
    const Fr* src[kMaxTables];
    Fr* dst[kMaxTables];
#pragma unroll
    for (int k = 0; k < kMaxTables; k++) {
      src[k] = cur < 0 ? ta.in[k] : (cur == 0 ? ta.b0[k] : ta.b1[k]);
      dst[k] = cur < 0 ? ta.b0[k] : (cur == 0 ? ta.b1[k] : ta.b0[k]);
    }
    Fr x = Fr::zero();
    if (n >= 4) {
      // fused round: bind src (n) into dst (n/2), evaluate the next round over dst (n/4 pairs)
      const size_t n_cur = n / 2, pairs = n_cur / 2, half_old = n / 2, half_cur = n_cur / 2;
      SplitEqDev sq{nullptr, nullptr, 0};
      if (ta.e_out) {  // split_eq_for(n_cur)
        int u = 0;
        while ((size_t(1) << u) < n_cur) u++;
        u = u > ta.eq_skip ? u - ta.eq_skip : 0;
        const int head_len = u > 0 ? u - 1 : 0, m = ta.n_tau / 2;
        const int out_bits = head_len < m ? head_len : m, in_bits = head_len - out_bits;
        sq.e_out = ta.e_out + ((size_t(1) << out_bits) - 1);
        sq.e_in = ta.e_in + ((size_t(1) << in_bits) - 1);
        sq.in_bits = in_bits;
        sq.out_mask = (1u << out_bits) - 1u;
      }
      const size_t n_chunks = (pairs + kTailChunk - 1) / kTailChunk;
      for (size_t c = cta; c < n_chunks; c += kTailCluster) {
#pragma unroll 1
        for (int it = tid; it < kTailChunk * 2 * kt; it += kTailThreads) {
          const int p = it % kTailChunk, h = (it / kTailChunk) & 1, k = it / (2 * kTailChunk);
          const size_t g = c * kTailChunk + p;
          if (g >= pairs) continue;
          const size_t idx = (PAIR == ZOLT_PAIR_ADJACENT) ? 2 * g + h : g + (h ? half_cur : 0);
          const Fr* s = src[k];
          const Fr lo = (PAIR == ZOLT_PAIR_ADJACENT) ? ldcg_fr(s + 2 * idx) : ldcg_fr(s + idx);
          const Fr hi = (PAIR == ZOLT_PAIR_ADJACENT) ? ldcg_fr(s + 2 * idx + 1) : ldcg_fr(s + idx + half_old);
          const Fr v = bind_pair_cold(lo, hi, r);
          st_fr(dst[k] + idx, v);
          if (k < a.k) stash[k][h][p] = v;
        }
        __syncthreads();
        {
          // warp w: raw sum (w & 3) of pairs (w >> 2) * 32 + lane of this chunk
          const int e = warp & 3, p = (warp >> 2) * 32 + lane;
          const size_t g = c * kTailChunk + p;
          if (g < pairs) x = fr_add_cold(x, small_term(a, e, stash, p, sq, g));
        }
        __syncthreads();
      }
    } else {
      // last variable: bind only
      if (cta == 0 && tid < kt) st_fr(dst[tid], bind_pair_cold(ldcg_fr(src[tid]), ldcg_fr(src[tid] + 1), r));
    }
    if (ta.dbg && cta == 0 && tid == 0) { unsigned long long tt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt)); ta.dbg[(expect - 1) * 4 + 1] = tt; }
    // ---- fold: warp -> CTA -> cluster ----
    x = warp_sum_cold(x);
    if (lane == 0) red[warp] = x;
    __syncthreads();
    if (tid < kRaw) {
      Fr s = red[tid];
#pragma unroll 1
      for (int w = tid + 4; w < kTailThreads / 32; w += 4) s = fr_add_cold(s, red[w]);
      st_fr(&ta.partials[cta * kRaw + tid], s);
    }
    cluster_sync_all();  // every CTA's partials and bound tables are written (release) and visible (acquire)
    if (ta.dbg && cta == 0 && tid == 0) { unsigned long long tt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt)); ta.dbg[(expect - 1) * 4 + 2] = tt; }
    if (cta == 0 && n >= 4) {
      // thread e folds raw sum e over the CTAs; thread 0 completes the polynomial and publishes it
      if (tid < kRaw) {
        Fr s = Fr::zero();
#pragma unroll 1
        for (int b = 0; b < kTailCluster; b++) s = fr_add(s, ldcg_fr(&ta.partials[b * kRaw + tid]));
        red[tid] = s;
      }
      __syncthreads();
      if (tid == 0) {
        Fr tot[kRaw], ev[4], claim;
#pragma unroll 1
        for (int e = 0; e < kRaw; e++) tot[e] = red[e];
#pragma unroll
        for (int i = 0; i < 8; i++) claim.v[i] = claim_w[i];
        sc_finalize_inl(a.deg, tot, claim, ev);
        Fr* h_evals = reinterpret_cast<Fr*>(ta.mailbox);
#pragma unroll 1
        for (int e = 0; e < 4; e++) st_fr(&h_evals[e], ev[e]);
        __threadfence_system();
        mb32[kMbResultSeq / 4] = cmd_sh[1];
#pragma unroll 1
        for (int e = 0; e < 4; e++) {
          st_fr(&ta.d_evals[e], ev[e]);
          st_fr(&ta.d_raw[e], tot[e]);
        }
      }
    }
    if (ta.dbg && cta == 0 && tid == 0) { unsigned long long tt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt)); ta.dbg[(expect - 1) * 4 + 3] = tt; }
    cluster_sync_all();  // partials may be overwritten, the next round may read dst and d_evals, cmd_sh may be rewritten
    cur = cur < 0 ? 0 : (cur ^ 1);
    n /= 2;
    expect++;
  }
  if (cta == 0 && tid == 0) {
    __threadfence_system();
    mb32[kMbTailStatus / 4] = status;
    __threadfence_system();
  }
  cluster_sync_all();  // no CTA leaves while another may still address its shared memory
}

// last variable (n == 2): bind every table's single pair, nothing left to evaluate
__global__ void bind_last_kernel(const __grid_constant__ ScPtrs t, int kt, const __grid_constant__ Fr r) {
  const int k = threadIdx.x;
  if (k < kt) st_fr(t.dst[k], bind_pair_cold(ld_fr(t.src[k]), ld_fr(t.src[k] + 1), r));
}

// Sharded runs: fold the all-gathered raw sums of every rank (rows x kRaw) and finish the round.
__global__ void sc_finish_rows_kernel(const Fr* rows, size_t n_rows, ScOut out, Fr r) {
  if (threadIdx.x != 0) return;
  Fr raw[kRaw], prev[4], ev[4];
  for (int c = 0; c < kRaw; c++) {
    Fr acc = Fr::zero();
    for (size_t i = 0; i < n_rows; i++) acc = fr_add(acc, ld_fr(rows + i * kRaw + c));
    raw[c] = acc;
  }
  Fr claim = Fr::zero();
  if (out.claim_src == CLAIM_PREV) {
    for (int e = 0; e < 4; e++) prev[e] = ld_fr(&out.d_evals[e]);
    claim = newton_eval4(prev, r);
  } else if (out.claim_src == CLAIM_HOST) {
    claim = ld_fr(out.d_claim);
  }
  sc_finalize(out.deg, out.claim_src != CLAIM_NONE, raw, claim, ev);
  for (int e = 0; e < 4; e++) {
    st_fr(&out.d_evals[e], ev[e]);
    st_fr(&out.h_evals[e], ev[e]);
  }
  __threadfence_system();
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
__global__ void __launch_bounds__(kScThreads) sum_kernel(const Fr* src, size_t n, ScOut out) {
  __shared__ Fr smem[32];
  Fr acc[1] = {Fr::zero()};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    acc[0] = fr_add(acc[0], ld_fr(src + i));
  grid_finish<1>(acc, smem, out, Fr::zero());
}

// BatchOps.innerProduct (src/field/mod.zig:1164-1280): sum_i a[i] * b[i], unreduced 512-bit accumulation per thread
__global__ void __launch_bounds__(kScThreads) inner_product_kernel(const Fr* a, const Fr* b, size_t n, ScOut out) {
  __shared__ Fr smem[32];
  uint32_t w[16];
  wide_zero(w);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const Fr x = ld_fr(a + i), y = ld_fr(b + i);
    wide_acc(w, x, y);
  }
  Fr acc[1] = {wide_collapse(w)};
  grid_finish<1>(acc, smem, out, Fr::zero());
}

// x (10 limbs, below 2^288) -> canonical x mod r; once per thread
struct Limbs10 { uint32_t v[10]; };
__device__ __noinline__ Fr lasso_reduce_288_cold(Limbs10 x) { return fp_barrett_289<FrParams>(x.v); }
__device__ __forceinline__ Fr lasso_reduce_288(const uint32_t* x) {
  Limbs10 t;
#pragma unroll
  for (int k = 0; k < 10; k++) t.v[k] = x[k];
  return lasso_reduce_288_cold(t);
}

// LassoProver address rounds (src/zkvm/lasso/prover.zig:265-310 computeAddressRoundPoly, :380-412 receiveChallenge): one pass
// that applies the PREVIOUS round's challenge -- eq_evals[j] *= bit(idx_j, bind_bit) ? r : 1 - r -- and accumulates this round's
// partition sums sum_b = sum over { j : bit(idx_j, sum_bit) = b } of eq_evals[j].  The reference's claim after the bind is the
// sum of the updated table, i.e. sum_0 + sum_1 of this very pass.  idx: 128-bit little-endian lookup indices.
__global__ void __launch_bounds__(kScThreads)
lasso_addr_kernel(Fr* __restrict__ eq, const uint4* __restrict__ idx, size_t n, int bind_bit, const __grid_constant__ FixedMulTable rt,
                  int sum_bit, ScOut out) {
  __shared__ Fr smem[64];
  // per-thread sums as plain 288-bit integers (canonical entries, at most 2^23 of them per thread: no overflow), reduced
  // once at the end: 9 carry-chain adds per entry and sum instead of an add and a conditional subtraction.
  // tot = all entries, one = the entries whose index has bit sum_bit set; sum_0 = tot - one.
  uint32_t tot[10], one[10];
#pragma unroll
  for (int k = 0; k < 10; k++) tot[k] = one[k] = 0;
  // two entries per iteration, both loaded before either is used: the pass is bound by memory latency, not by arithmetic
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  auto step = [&](size_t j, const uint4& w, Fr e) {
    const uint32_t limb[4] = {w.x, w.y, w.z, w.w};
    if (bind_bit >= 0) {
      // the multiplier is the same for every entry of a launch: one product through the table of r (field.cuh
      // fp_fixed_mul_add, 80 wide multiplies instead of 136), and e (1 - r) = e - e r costs a subtraction
      const Fr er = fp_fixed_mul_add<FrParams>(Fr::zero(), e.v, rt);
      e = ((limb[bind_bit >> 5] >> (bind_bit & 31)) & 1u) ? er : fr_sub(e, er);
      st_fr(eq + j, e);
    }
    const uint32_t mask = sum_bit >= 0 ? 0u - ((limb[sum_bit >> 5] >> (sum_bit & 31)) & 1u) : 0u;
    tot[0] = ptx::add_cc(tot[0], e.v[0]);
#pragma unroll
    for (int k = 1; k < 8; k++) tot[k] = ptx::addc_cc(tot[k], e.v[k]);
    tot[8] = ptx::addc(tot[8], 0u);
    one[0] = ptx::add_cc(one[0], e.v[0] & mask);
#pragma unroll
    for (int k = 1; k < 8; k++) one[k] = ptx::addc_cc(one[k], e.v[k] & mask);
    one[8] = ptx::addc(one[8], 0u);
  };
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; j + stride < n; j += 2 * stride) {
    const uint4 w0 = __ldg(idx + j), w1 = __ldg(idx + j + stride);
    const Fr e0 = eq[j], e1 = eq[j + stride];
    step(j, w0, e0);
    step(j + stride, w1, e1);
  }
  if (j < n) step(j, __ldg(idx + j), eq[j]);
  const Fr s_one = lasso_reduce_288(one), s_tot = lasso_reduce_288(tot);
  Fr acc[2] = {fr_sub_cold(s_tot, s_one), s_one};
  grid_finish<2>(acc, smem, out, Fr::zero());
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

// EqPlusOnePolynomial.mle(r, j) for every j in {0,1}^n (src/poly/mod.zig:387-421, tabulated by stage3_prover.zig:1878-1895 with
// one O(n^2) mle per entry).  For a boolean point only the term k = (number of trailing zeros of j) survives -- the low k bits
// of y must be 0 and bit k must be 1 -- so entry j is  prod_{i<k} r_i' * (1 - r_k') * prod_{i>k} (j_i ? r_i' : 1 - r_i'),
// r_i' = r[n-1-i] (bit 0 of the index pairs with the LAST coordinate); entry 0 is zero.  The vanishing terms are exact zeros
// in the reference's sum, so the field elements are the same.
__global__ void __launch_bounds__(kScThreads) eq_plus_one_kernel(Fr* __restrict__ out, const Fr* __restrict__ r, int n) {
  const size_t size = size_t(1) << n;
  for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < size; j += (size_t)gridDim.x * blockDim.x) {
    if (j == 0) {
      st_fr(out, Fr::zero());
      continue;
    }
    const int k = __ffsll((long long)j) - 1;
    Fr acc = fr_sub(Fr::one(), r[n - 1 - k]);
    for (int i = 0; i < k; i++) acc = fr_mul(acc, r[n - 1 - i]);
    for (int i = k + 1; i < n; i++) {
      const Fr x = r[n - 1 - i];
      acc = fr_mul(acc, ((j >> i) & 1) ? x : fr_sub(Fr::one(), x));
    }
    st_fr(out + j, acc);
  }
}

// out[x_lo] = scale * sum_{x_hi} in[x_lo + x_hi * n_lo] * w[x_hi]: the Q buffers of the prefix-suffix provers
// (stage3_prover.zig:1053-1112 ShiftPrefixSuffixProver.init, :2262-2300 RegistersPrefixSuffixProver.init), an O(T) weighted
// fold of a trace column over the HIGH half of the cycle index.  A warp reads 32 consecutive x_lo (1 KB) of one row, the
// rows are dealt to the warps of the block and to gridDim.y row slices; unreduced 512-bit sums per thread.
constexpr int kFoldWarps = 8;
__global__ void __launch_bounds__(32 * kFoldWarps)
fold_high_kernel(const Fr* __restrict__ in, const Fr* __restrict__ w, size_t n_lo, size_t n_hi, Fr* __restrict__ partial) {
  __shared__ Fr smem[kFoldWarps][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t x_lo = (size_t)blockIdx.x * 32 + lane;
  uint32_t acc[16];
  wide_zero(acc);
  if (x_lo < n_lo) {
    for (size_t x_hi = (size_t)blockIdx.y * kFoldWarps + warp; x_hi < n_hi; x_hi += (size_t)gridDim.y * kFoldWarps) {
      const Fr v = ld_fr(in + x_lo + x_hi * n_lo), s = ld_fr(w + x_hi);
      wide_acc(acc, v, s);
    }
  }
  smem[warp][lane] = wide_collapse(acc);
  __syncthreads();
  if (warp == 0 && x_lo < n_lo) {
    Fr t = smem[0][lane];
#pragma unroll 1
    for (int k = 1; k < kFoldWarps; k++) t = fr_add_cold(t, smem[k][lane]);
    st_fr(partial + (size_t)blockIdx.y * n_lo + x_lo, t);
  }
}
__global__ void __launch_bounds__(kScThreads)
fold_finish_kernel(const Fr* __restrict__ partial, int slices, size_t n_lo, Fr* __restrict__ out, const Fr* __restrict__ scale) {
  const size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (x >= n_lo) return;
  Fr t = ld_fr(partial + x);
#pragma unroll 1
  for (int k = 1; k < slices; k++) t = fr_add_cold(t, ld_fr(partial + (size_t)k * n_lo + x));
  if (scale) t = fr_mul_cold(t, ld_fr(scale));
  st_fr(out + x, t);
}

// Element-wise construction of device-resident tables, one pass over HBM each.
template <int OP>
__global__ void __launch_bounds__(kScThreads)
vec_kernel(Fr* __restrict__ out, const Fr* __restrict__ a, const Fr* __restrict__ b, const Fr* __restrict__ c, const Fr* __restrict__ e,
           const __grid_constant__ FixedMulTable st, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    Fr v;
    if constexpr (OP == ZOLT_VEC_MUL) {
      v = fr_mul(ld_fr(a + i), ld_fr(b + i));
    } else if constexpr (OP == ZOLT_VEC_EQ_AB_MINUS_C) {
      // combined[i] = eq[i] * (Az[i] * Bz[i] - Cz[i])  (src/zkvm/r1cs/jolt_r1cs.zig:352-360)
      v = fr_mul(ld_fr(e + i), fr_sub(fr_mul(ld_fr(a + i), ld_fr(b + i)), ld_fr(c + i)));
    } else if constexpr (OP == ZOLT_VEC_AFFINE_INDEX) {
      // out[i] = s0 + s1 * i: UnmapPolynomial (start_address + 8 k, src/zkvm/ram/raf_checking.zig:206-209) as a dense table;
      // `a` carries the two scalars, i enters Montgomery form through one product by R^2
      Fr idx = Fr::zero();
      idx.v[0] = (uint32_t)i;
      idx.v[1] = (uint32_t)((unsigned long long)i >> 32);
      v = fr_add(a[0], fr_mul(a[1], fr_mul(Fr::r2(), idx)));
    } else if constexpr (OP == ZOLT_VEC_LIN2) {
      // s0 a + s1 b with the scalars in c[0], c[1]: ra = gamma rs1_ra + gamma^2 rs2_ra (stage4_gruen_prover.zig:211, :224)
      v = fp_mul2_add<FrParams>(c[0], ld_fr(a + i), c[1], ld_fr(b + i));
    } else if constexpr (OP == ZOLT_VEC_STAGE4_TERM) {
      // eq (ra val + wa (val + inc)), inc in `out` (computeInputClaim, stage4_gruen_prover.zig:385-430)
      const Fr val = ld_fr(b + i);
      v = fr_mul(ld_fr(e + i), fp_mul2_add<FrParams>(ld_fr(a + i), val, ld_fr(c + i), fr_add(val, out[i])));
    } else {
      // out[i] += s * a[i]: the gamma-power combination of batchOpen (src/poly/commitment/mod.zig:648-656)
      const Fr x = ld_fr(a + i);
      v = fp_fixed_mul_add<FrParams>(out[i], x.v, st);
    }
    st_fr(out + i, v);
  }
}

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

inline unsigned grid_for(const zolt_ctx* ctx, size_t work_items, int per_sm, int threads = kScThreads) {
  size_t blocks = (work_items + threads - 1) / threads;
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
  bool evals_valid = false;   // a round kernel for the current tables has been launched
  bool evals_final = false;   // ... and its raw sums have been turned into d_evals (always true unless sharded)
  bool sharded = false;       // multi-GPU: kernels stop at the raw sums, zolt_sc_finish_partials completes the round
  bool have_claim = false;    // zolt_sc_set_claim: round 0 derives p(1) from it
  int pending_claim_src = 0;  // how the launched-but-unfinished round (sharded) gets p(1)
  Fr last_r;                  // challenge of the most recent bind
  Fr* d_raw = nullptr;
  Fr* d_claim = nullptr;
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
  unsigned wait_seq = 0;         // mailbox word that marks the evaluations of the latest launch as delivered
  zolt_peer_box* box = nullptr;  // zolt_sc_set_peers: rounds exchange their sums through the NVLink mailboxes
  // persistent tail kernel (sc_tail_kernel): resident from the first small round to the last bind
  cudaStream_t tail_stream = nullptr;
  cudaEvent_t tail_ev = nullptr;
  bool tail_active = false;     // the kernel has been launched and not yet joined
  unsigned tail_cmds = 0;       // commands posted to the running kernel
  int tail_prof = -1;
  unsigned tail_bind_seq = 0;   // number of the last BIND command posted (0: none) -- compared with the kernel's acknowledgement
  bool tail_pending = false;    // a bind command has been posted whose evaluations have not been read yet
  int tail_prev_cur = 0;        // state before that command (to redo it through a regular launch if the kernel timed out)
  size_t tail_prev_n = 0;
  // tensor maps of the next fused launch, encoded ahead of time (launch_round)
  TmapArray tm_next{};
  size_t tm_next_n = 0;
  const Fr* tm_next_src0 = nullptr;
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
    case ZOLT_SC_STAGE4_GRUEN: return 4;
    case ZOLT_SC_STAGE4_DENSE: return 5;
    case ZOLT_SC_EQ_DOT2: return 5;
    default: return -1;
  }
}
int combiner_ne(int comb) {
  if (comb == ZOLT_SC_STAGE4_DENSE || comb == ZOLT_SC_EQ_DOT2) return 3;
  return (comb == ZOLT_SC_SUM || comb == ZOLT_SC_GRUEN_AB || comb == ZOLT_SC_GRUEN_AB_MINUS_C || comb == ZOLT_SC_STAGE4_GRUEN) ? 2 : 4;
}
bool combiner_gruen(int comb) { return comb == ZOLT_SC_GRUEN_AB || comb == ZOLT_SC_GRUEN_AB_MINUS_C || comb == ZOLT_SC_STAGE4_GRUEN; }
// the combiners written for the Stage 3 / Stage 4 provers: adjacent pairs only, never sharded
bool combiner_stage4(int comb) { return comb == ZOLT_SC_STAGE4_GRUEN || comb == ZOLT_SC_STAGE4_DENSE || comb == ZOLT_SC_EQ_DOT2; }

SplitEqDev split_eq_for(const zolt_sc* sc, size_t n_cur) {
  // getWindowEqTables with window_size 1 (src/poly/split_eq.zig:311-343); u = unbound variables.
  SplitEqDev s{nullptr, nullptr, 0};
  if (!combiner_gruen(sc->combiner)) return s;
  int u = 0;
  while ((size_t(1) << u) < n_cur) u++;
  const int skip = sc->log_n0 - sc->n_tau;  // Stage 4: the weights cover the cycle bits, not the row k above them
  u = u > skip ? u - skip : 0;
  const int head_len = u > 0 ? u - 1 : 0;
  const int m = sc->n_tau / 2;
  const int out_bits = head_len < m ? head_len : m;
  const int in_bits = head_len - out_bits;
  s.e_out = sc->d_e_out + ((size_t(1) << out_bits) - 1);
  s.e_in = sc->d_e_in + ((size_t(1) << in_bits) - 1);
  s.in_bits = in_bits;
  s.out_mask = (1u << out_bits) - 1u;
  return s;
}

int combiner_deg(int comb) {
  switch (comb) {
    case ZOLT_SC_PROD2: return 2;
    case ZOLT_SC_PROD3:
    case ZOLT_SC_EQ_A_BMC:
    case ZOLT_SC_EQ_AB_MINUS_C: return 3;
    default: return 0;
  }
}

constexpr int kBlocksPerSm = kFusedBlocksPerSm;  // upper bound, sizes the partials buffer

// Table of the bind multiplier for fp_bind_fixed: row 0 is r as a plain integer (one Montgomery product by 1), row
// i+1 = row i * 2^32 mod p by 32 modular doublings on 64-bit limbs (cheaper on the host than the PTX emulation).
FixedMulTable fixed_mul_table_host(const Fr& r) {
  static const uint64_t P[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
  Fr one_raw = Fr::zero();
  one_raw.v[0] = 1;
  const Fr r_std = fp_mul<FrParams>(r, one_raw);
  uint64_t c[4];
  for (int j = 0; j < 4; j++) c[j] = (uint64_t)r_std.v[2 * j] | ((uint64_t)r_std.v[2 * j + 1] << 32);
  auto doublings = [&](int count) {
    for (int d = 0; d < count; d++) {
      // c = 2c mod p  (c < p < 2^254: no overflow)
      c[3] = (c[3] << 1) | (c[2] >> 63);
      c[2] = (c[2] << 1) | (c[1] >> 63);
      c[1] = (c[1] << 1) | (c[0] >> 63);
      c[0] <<= 1;
      uint64_t s[4];
      unsigned __int128 bw = 0;
      for (int j = 0; j < 4; j++) {
        const unsigned __int128 x = (unsigned __int128)c[j] - P[j] - (uint64_t)bw;
        s[j] = (uint64_t)x;
        bw = (x >> 64) & 1;
      }
      if (!bw)
        for (int j = 0; j < 4; j++) c[j] = s[j];
    }
  };
#if defined(ZOLT_BIND_MONT1)
  doublings(32);  // one reduction row: rows scaled by 2^32 (field.cuh fp_fixed_mul_add_mont1)
#elif !defined(ZOLT_BIND_BARRETT)
  doublings(64);  // the two-limb Montgomery form of the bind takes the rows scaled by 2^64 (field.cuh fp_fixed_mul_add_mont2)
#endif
  FixedMulTable t;
  for (int i = 0; i < 8; i++) {
    for (int j = 0; j < 4; j++) {
      t.c[i][2 * j] = (uint32_t)c[j];
      t.c[i][2 * j + 1] = (uint32_t)(c[j] >> 32);
    }
    doublings(32);
  }
  return t;
}

// FP64 form of the same table (fp_bind_fixed_f64): row l = limbs of r_std * 2^(16 l) mod p as doubles.  Rows 2i are the
// rows of the integer table; rows 2i+1 are 16 modular doublings further.
void fixed_mul_table_f64_host(const Fr& r, FixedMulTableF64* out) {
  static const uint64_t P[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
  Fr one_raw = Fr::zero();
  one_raw.v[0] = 1;
  const Fr r_std = fp_mul<FrParams>(r, one_raw);
  uint64_t c[4];
  for (int j = 0; j < 4; j++) c[j] = (uint64_t)r_std.v[2 * j] | ((uint64_t)r_std.v[2 * j + 1] << 32);
  for (int l = 0; l < 16; l++) {
    for (int j = 0; j < 4; j++) {
      out->c[l][j] = F64Pair{(double)(uint32_t)c[j], (double)(uint32_t)(c[j] >> 32)};
    }
    for (int d = 0; d < 16; d++) {
      c[3] = (c[3] << 1) | (c[2] >> 63);
      c[2] = (c[2] << 1) | (c[1] >> 63);
      c[1] = (c[1] << 1) | (c[0] >> 63);
      c[0] <<= 1;
      uint64_t s[4];
      unsigned __int128 bw = 0;
      for (int j = 0; j < 4; j++) {
        const unsigned __int128 x = (unsigned __int128)c[j] - P[j] - (uint64_t)bw;
        s[j] = (uint64_t)x;
        bw = (x >> 64) & 1;
      }
      if (!bw)
        for (int j = 0; j < 4; j++) c[j] = s[j];
    }
  }
}

// 2-D tensor map over a table viewed as rows of `row_bytes` (64 or 128): box = 32 rows, swizzle = the row width.
// The encoder is a driver entry point; it is fetched through the runtime so the library links against cudart only.
int encode_rows_tmap(CUtensorMap* out, const void* base, size_t rows, unsigned row_bytes) {
  static PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
  if (!enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    ZOLT_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) {
      fprintf(stderr, "[zolt_gpu] cuTensorMapEncodeTiled is not available from this driver\n");
      return ZOLT_ERR_CUDA;
    }
    enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  const cuuint64_t gdim[2] = {row_bytes / 4, rows ? rows : 1};
  const cuuint64_t gstride[1] = {row_bytes};
  const cuuint32_t box[2] = {row_bytes / 4, 32};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult rc = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    fprintf(stderr, "[zolt_gpu] cuTensorMapEncodeTiled failed (%d) for %zu rows of %u bytes at %p\n", (int)rc, rows, row_bytes, base);
    return ZOLT_ERR_CUDA;
  }
  return ZOLT_OK;
}

template <int COMB, int PAIR>
int launch_round(zolt_sc* sc, bool fused, const Fr& r) {
  constexpr int COMBK = CombTraits<COMB>::K;
  ScPtrs t{};
  const int total = sc->k + sc->n_extra;
  for (int i = 0; i < total; i++) {
    t.src[i] = sc->cur < 0 ? sc->tab_in[i] : sc->buf[sc->cur][i];
    t.dst[i] = sc->buf[sc->cur < 0 ? 0 : (sc->cur ^ 1)][i];
  }
  const size_t n_cur = fused ? sc->n / 2 : sc->n;
  const unsigned grid = grid_for(sc->ctx, n_cur / 2, fused ? kFusedBlocksPerSm : kEvalBlocksPerSm, kRoundThreads);
  SplitEqDev sq = split_eq_for(sc, n_cur);
  ScOut out{};
  out.partials = sc->d_partials;
  out.ticket = sc->d_ticket;
  out.next_chunk = sc->d_ticket + 32;  // its own 128-byte line
  out.d_raw = sc->d_raw;
  out.d_evals = sc->d_evals;
  out.h_evals = sc->h_evals_dev;
  out.d_claim = sc->d_claim;
  out.d_claim_cur = sc->d_claim + 1;
  out.deg = combiner_deg(sc->combiner);
  out.finalize = sc->sharded ? 0 : 1;
  out.peer = peer_next(sc->sharded ? nullptr : sc->box);
  if (out.finalize) {
    if (++sc->ctx->mail_seq == 0) ++sc->ctx->mail_seq;
    out.seq = sc->wait_seq = sc->ctx->mail_seq;
  }
  out.dbg = sc->ctx->sc_round_dbg;
  const int mode = fused ? 2 : (sc->have_claim && sc->cur < 0 && out.deg != 0 ? 1 : 0);
  out.claim_src = mode == 2 ? CLAIM_PREV : (mode == 1 ? CLAIM_HOST : CLAIM_NONE);
  sc->pending_claim_src = out.claim_src;
  // shared memory: the layout of this pairing (see the kernel) + the per-warp mbarriers + slack for 1024-byte alignment
  const size_t kWarps = kRoundThreads / 32;
  const size_t smem_unit = (size_t)(fused ? kFusedScratchUnits : kEvalScratchUnits) * kScrStride * sizeof(uint4);
  const size_t smem_tma = (fused ? kWarps * (32 * 128 + 4 * (size_t)COMBK * 32 * 16) + kWarps * 8 : kWarps * 2 * (size_t)COMBK * 32 * 64 + kWarps * 16) + 1024;
  const size_t kDynSmem = PAIR == ZOLT_PAIR_ADJACENT ? smem_tma : smem_unit;
  TmapArray tm{};
  if (PAIR == ZOLT_PAIR_ADJACENT && !(n_cur / 2 <= (size_t)sc->ctx->sc_small_pairs)) {
    // one row = what one pair reads from a table: 4 source elements (fused) or 2 (evaluation only)
    const unsigned row_bytes = fused ? 128u : 64u;
    if (fused && sc->tm_next_n == sc->n && sc->tm_next_src0 == t.src[0]) {
      tm = sc->tm_next;  // encoded while the previous round was running (below)
    } else {
      for (int i = 0; i < (fused ? total : sc->k); i++) ZOLT_TRY(encode_rows_tmap(&tm.m[i], t.src[i], sc->n * sizeof(Fr) / row_bytes, row_bytes));
    }
  }
  const bool f64 = mode == 2 && PAIR == ZOLT_PAIR_ADJACENT && (sc->ctx->sc_bind_f64 & 0xff) != 0;  // sc_bind_f64: mask of the tables whose binds take the FP64 pipe
  {
    // cudaFuncSetAttribute applies to the CURRENT device only: remember it per (instantiation, device), so a process
    // that drives several GPUs (zolt_gpu_init_multi) opts every one of them in to the 64 KB of dynamic shared memory
    static std::atomic<unsigned long long> attr_devs[4];  // per (COMB, PAIR) instantiation of this function; bit = device
    const int slot = mode == 2 ? (f64 ? 3 : 2) : mode;
    const unsigned long long bit = 1ull << (sc->ctx->device & 63);
    if (!(attr_devs[slot].load(std::memory_order_acquire) & bit)) {
      cudaError_t e;
      if (slot == 3) e = cudaFuncSetAttribute(sc_round_kernel<COMB, PAIR, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDynSmem);
      else if (slot == 2) e = cudaFuncSetAttribute(sc_round_kernel<COMB, PAIR, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDynSmem);
      else if (slot == 1) e = cudaFuncSetAttribute(sc_round_kernel<COMB, PAIR, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDynSmem);
      else e = cudaFuncSetAttribute(sc_round_kernel<COMB, PAIR, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDynSmem);
      if (e != cudaSuccess) {
        fprintf(stderr, "[zolt_gpu] cudaFuncSetAttribute(dynamic shared memory) failed on device %d: %s\n", sc->ctx->device,
                cudaGetErrorString(e));
        return ZOLT_ERR_CUDA;
      }
      attr_devs[slot].fetch_or(bit, std::memory_order_release);
    }
  }
  const BindTableI32 rt0{};
  const int ps = prof_begin(sc->ctx, fused ? PROF_SC_FUSED : PROF_SC_EVAL);
  if (n_cur / 2 <= (size_t)sc->ctx->sc_small_pairs) {
    // latency-bound sizes: one element / one term per thread (sc_small_kernel)
    SmallArgs a{sc->combiner, sc->k, sc->n_extra, out.deg, fused ? 1 : 0, mode == 0 ? 1 : 0};
    size_t chunks = (n_cur / 2 + kSmallPairs - 1) / kSmallPairs;
    if (chunks > sc->max_grid) chunks = sc->max_grid;
    if (chunks == 0) chunks = 1;
    sc_small_kernel<PAIR><<<(unsigned)chunks, kSmallThreads, 0, sc->ctx->stream>>>(t, a, sc->n, r, sq, out);
  } else if (mode == 2) {
    if (f64) {
      BindTables rt{};  // O(1) host glue per launch
      fixed_mul_table_f64_host(r, &rt.f64[0]);
      rt.f64[1] = rt.f64[0];
      rt.i32 = fixed_mul_table_host(r);
      rt.f64_mask = (unsigned)sc->ctx->sc_bind_f64 & 0xffu;
      sc_round_kernel<COMB, PAIR, 2, 1><<<grid, kRoundThreads, kDynSmem, sc->ctx->stream>>>(t, sc->n_extra, sc->n, r, sq, out, rt, tm);
    } else {
      BindTableI32 rt{};
      rt.i32 = fixed_mul_table_host(r);
      sc_round_kernel<COMB, PAIR, 2, 0><<<grid, kRoundThreads, kDynSmem, sc->ctx->stream>>>(t, sc->n_extra, sc->n, r, sq, out, rt, tm);
    }
  } else if (mode == 1) {
    sc_round_kernel<COMB, PAIR, 1><<<grid, kRoundThreads, kDynSmem, sc->ctx->stream>>>(t, sc->n_extra, sc->n, r, sq, out, rt0, tm);
  } else {
    sc_round_kernel<COMB, PAIR, 0><<<grid, kRoundThreads, kDynSmem, sc->ctx->stream>>>(t, sc->n_extra, sc->n, r, sq, out, rt0, tm);
  }
  prof_end(sc->ctx, ps);
  sc->ctx->launches++;
  sc->evals_final = !sc->sharded;
  // The NEXT launch is a fused pass over the tables this one evaluates (n_cur entries, at t.src when this launch did not
  // bind, else at t.dst): encode its tensor maps now, while the GPU works, instead of between the challenge and the launch.
  sc->tm_next_n = 0;
  if (PAIR == ZOLT_PAIR_ADJACENT && n_cur >= 4 && n_cur / 4 > (size_t)sc->ctx->sc_small_pairs) {
    bool ok = true;
    for (int i = 0; i < total && ok; i++)
      ok = encode_rows_tmap(&sc->tm_next.m[i], fused ? t.dst[i] : t.src[i], n_cur * sizeof(Fr) / 128u, 128u) == ZOLT_OK;
    if (ok) {
      sc->tm_next_n = n_cur;
      sc->tm_next_src0 = fused ? t.dst[0] : t.src[0];
    }
  }
  return ZOLT_OK;
}

template <int COMB>
int launch_round_p(zolt_sc* sc, bool fused, const Fr& r) {
  if (sc->pairing == ZOLT_PAIR_ADJACENT) return launch_round<COMB, ZOLT_PAIR_ADJACENT>(sc, fused, r);
  return launch_round<COMB, ZOLT_PAIR_HALF>(sc, fused, r);
}

int launch_round_any(zolt_sc* sc, bool fused, const Fr& r) {
  switch (sc->combiner) {
    case ZOLT_SC_SUM: return launch_round_p<ZOLT_SC_SUM>(sc, fused, r);
    case ZOLT_SC_PROD2: return launch_round_p<ZOLT_SC_PROD2>(sc, fused, r);
    case ZOLT_SC_PROD3: return launch_round_p<ZOLT_SC_PROD3>(sc, fused, r);
    case ZOLT_SC_EQ_A_BMC: return launch_round_p<ZOLT_SC_EQ_A_BMC>(sc, fused, r);
    case ZOLT_SC_EQ_AB_MINUS_C: return launch_round_p<ZOLT_SC_EQ_AB_MINUS_C>(sc, fused, r);
    case ZOLT_SC_GRUEN_AB: return launch_round_p<ZOLT_SC_GRUEN_AB>(sc, fused, r);
    case ZOLT_SC_GRUEN_AB_MINUS_C: return launch_round_p<ZOLT_SC_GRUEN_AB_MINUS_C>(sc, fused, r);
    // the Stage-4 combiners exist for adjacent pairs only (sc_alloc refuses the other pairing)
    case ZOLT_SC_STAGE4_GRUEN: return launch_round<ZOLT_SC_STAGE4_GRUEN, ZOLT_PAIR_ADJACENT>(sc, fused, r);
    case ZOLT_SC_STAGE4_DENSE: return launch_round<ZOLT_SC_STAGE4_DENSE, ZOLT_PAIR_ADJACENT>(sc, fused, r);
    default: return launch_round<ZOLT_SC_EQ_DOT2, ZOLT_PAIR_ADJACENT>(sc, fused, r);
  }
}

int sc_alloc(zolt_ctx* ctx, int combiner, int pairing, int log_n, int n_tables, int n_extra, zolt_sc** out) {
  if (!ctx || !out) return ZOLT_ERR_ARG;
  *out = nullptr;
  const int k = combiner_k(combiner);
  if (k < 0 || n_tables != k) return ZOLT_ERR_ARG;
  if (pairing != ZOLT_PAIR_ADJACENT && pairing != ZOLT_PAIR_HALF) return ZOLT_ERR_ARG;
  if ((combiner_gruen(combiner) || combiner_stage4(combiner)) && pairing != ZOLT_PAIR_ADJACENT) return ZOLT_ERR_ARG;  // Gruen provers bind low-to-high only
  if (log_n < 0 || log_n > 40 || n_extra < 0 || k + n_extra > kMaxTables) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  zolt_sc* sc = new (std::nothrow) zolt_sc();
  if (!sc) return ZOLT_ERR_ARG;
  OnFail undo([&] { zolt_sc_free(sc); });  // a half-built instance: zolt_sc_free skips what is still null
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
  sc->max_grid = (unsigned)ctx->num_sms * kBlocksPerSm;
  ZOLT_CUDA_TRY(pool_alloc(ctx, (size_t)sc->max_grid * 4 * sizeof(Fr), (void**)&sc->d_partials));
  ZOLT_CUDA_TRY(pool_alloc(ctx, 4 * sizeof(Fr), (void**)&sc->d_raw));
  ZOLT_CUDA_TRY(pool_alloc(ctx, 4 * sizeof(Fr), (void**)&sc->d_claim));
  ZOLT_CUDA_TRY(pool_alloc(ctx, 256, (void**)&sc->d_ticket));
  ZOLT_CUDA_TRY(cudaMemsetAsync(sc->d_ticket, 0, 256, ctx->stream));
  ZOLT_CUDA_TRY(pool_alloc(ctx, 4 * sizeof(Fr), (void**)&sc->d_evals));
  ZOLT_CUDA_TRY(pinned_mailbox_alloc(ctx, (void**)&sc->h_evals, (void**)&sc->h_evals_dev));
  undo.dismiss();
  *out = sc;
  return ZOLT_OK;
}

}  // namespace

namespace {

inline volatile unsigned* mb_word(zolt_sc* sc, int byte_off) {
  return reinterpret_cast<volatile unsigned*>(reinterpret_cast<unsigned char*>(sc->h_evals) + byte_off);
}

bool tail_eligible(const zolt_sc* sc) {
  return sc->ctx->sc_tail_pairs > 0 && !sc->sharded && !sc->box && sc->k <= 4 && sc->k + sc->n_extra <= kMaxTables;  // 5-table combiners (Stage-4 dense): measured no faster in the tail than as small-round launches
}

int tail_launch(zolt_sc* sc) {
  zolt_ctx* ctx = sc->ctx;
  if (!sc->tail_stream) {
    if (!ctx->tail_streams_free.empty()) {
      sc->tail_stream = ctx->tail_streams_free.back();
      ctx->tail_streams_free.pop_back();
    } else {
      ZOLT_CUDA_TRY(cudaStreamCreateWithFlags(&sc->tail_stream, cudaStreamNonBlocking));
    }
    ZOLT_CUDA_TRY(cudaEventCreateWithFlags(&sc->tail_ev, cudaEventDisableTiming));
  }
  TailArgs ta{};
  const int total = sc->k + sc->n_extra;
  for (int i = 0; i < total; i++) {
    ta.in[i] = sc->tab_in[i];
    ta.b0[i] = sc->buf[0][i];
    ta.b1[i] = sc->buf[1][i];
  }
  ta.cur = sc->cur;
  ta.n = sc->n;
  ta.a = SmallArgs{sc->combiner, sc->k, sc->n_extra, combiner_deg(sc->combiner), 1, 0};
  ta.e_out = combiner_gruen(sc->combiner) ? sc->d_e_out : nullptr;
  ta.eq_skip = sc->log_n0 - sc->n_tau;
  ta.e_in = sc->d_e_in;
  ta.n_tau = sc->n_tau;
  ta.partials = sc->d_partials;
  ta.d_evals = sc->d_evals;
  ta.d_raw = sc->d_raw;
  ta.mailbox = reinterpret_cast<unsigned char*>(sc->h_evals_dev);
  ta.timeout_ns = (unsigned long long)(ctx->sc_tail_timeout_ms > 0 ? ctx->sc_tail_timeout_ms : 10000) * 1000000ull;
  ta.first_timeout_ns = ta.timeout_ns < 250000000ull ? ta.timeout_ns : 250000000ull;
  ta.dbg = ctx->sc_tail_dbg;
  memset(reinterpret_cast<unsigned char*>(sc->h_evals) + kMbCmd, 0, 16 * kCmdChunks);
  *mb_word(sc, kMbCmdAck) = 0;
  *mb_word(sc, kMbTailStatus) = TAIL_ST_RUNNING;
  __atomic_thread_fence(__ATOMIC_SEQ_CST);
  // the kernel continues from what the context's stream has produced so far (the tables, d_evals)
  ZOLT_CUDA_TRY(cudaEventRecord(sc->tail_ev, ctx->stream));
  ZOLT_CUDA_TRY(cudaStreamWaitEvent(sc->tail_stream, sc->tail_ev, 0));
  sc->tail_prof = prof_begin_on(ctx, PROF_SC_TAIL, sc->tail_stream);
  if (sc->pairing == ZOLT_PAIR_ADJACENT) sc_tail_kernel<ZOLT_PAIR_ADJACENT><<<kTailCluster, kTailThreads, 0, sc->tail_stream>>>(ta);
  else sc_tail_kernel<ZOLT_PAIR_HALF><<<kTailCluster, kTailThreads, 0, sc->tail_stream>>>(ta);
  ZOLT_CUDA_TRY(cudaGetLastError());
  prof_end_on(ctx, sc->tail_prof, sc->tail_stream);
  ctx->launches++;
  sc->tail_active = true;
  sc->tail_cmds = 0;
  sc->tail_bind_seq = 0;
  if (cudaStreamQuery(sc->tail_stream) == cudaSuccess) {
    // The kernel is over before its first command could be posted: kernel launches are being serialised (a profiler such
    // as ncu, CUDA_LAUNCH_BLOCKING), so a kernel that waits for the host can only time out.  Use one launch per round on
    // this context from now on.
    fprintf(stderr, "[zolt_gpu] kernel launches are serialised on device %d: resident sumcheck tail disabled for this context\n", ctx->device);
    ctx->sc_tail_pairs = 0;
    sc->tail_active = false;
    return ZOLT_ERR_STATE;
  }
  return ZOLT_OK;
}

void tail_post(zolt_sc* sc, unsigned op, const Fr* r, unsigned result_seq, const Fr* claim = nullptr) {
  // one command slot: the previous command must have been taken (acknowledged) before it is rewritten
  for (unsigned spins = 1; *mb_word(sc, kMbCmdAck) != sc->tail_cmds; spins++) {
    if ((spins & 0xfffu) == 0 && cudaStreamQuery(sc->tail_stream) != cudaErrorNotReady) break;  // the kernel is gone (timed out)
  }
  const unsigned seq = ++sc->tail_cmds;
  if (op == TAIL_OP_BIND) sc->tail_bind_seq = seq;
  unsigned w[18] = {op, result_seq};
  if (r)
    for (int i = 0; i < 8; i++) w[2 + i] = r->v[i];
  if (claim)
    for (int i = 0; i < 8; i++) w[10 + i] = claim->v[i];
  unsigned char* const base = reinterpret_cast<unsigned char*>(sc->h_evals) + kMbCmd;
  for (int c = 0; c < kCmdChunks; c++) {  // one aligned 16-byte store per chunk: a chunk is never seen half written
    alignas(16) const unsigned chunk[4] = {seq, w[3 * c], w[3 * c + 1], w[3 * c + 2]};
    _mm_store_si128(reinterpret_cast<__m128i*>(base + 16 * c), _mm_load_si128(reinterpret_cast<const __m128i*>(chunk)));
  }
  _mm_sfence();
}

// p(r) for the cubic through (0, e0) .. (3, e3), on the host (Newton forward differences; exact field arithmetic, so the
// same canonical value as any other evaluation formula -- evaluateCubicAtPointGeneric, batched_sumcheck.zig:437-457)
Fr host_cubic_at(const Fr* e, const Fr& r) {
  using P = FrParams;
  const Fr d1 = fp_sub<P>(e[1], e[0]), d1b = fp_sub<P>(e[2], e[1]), d1c = fp_sub<P>(e[3], e[2]);
  const Fr d2 = fp_sub<P>(d1b, d1), d2b = fp_sub<P>(d1c, d1b), d3 = fp_sub<P>(d2b, d2);
  Fr inv2, inv6;
  for (int i = 0; i < 8; i++) {
    inv2.v[i] = P::inv2(i);
    inv6.v[i] = P::inv6(i);
  }
  const Fr rm1 = fp_sub<P>(r, Fr::one()), rm2 = fp_sub<P>(rm1, Fr::one());
  const Fr rr1 = fp_mul<P>(r, rm1);
  const Fr c2 = fp_mul<P>(rr1, inv2), c3 = fp_mul<P>(fp_mul<P>(rr1, rm2), inv6);
  Fr acc = fp_add<P>(e[0], fp_mul<P>(r, d1));
  acc = fp_add<P>(acc, fp_mul<P>(c2, d2));
  return fp_add<P>(acc, fp_mul<P>(c3, d3));
}

// Join the tail kernel (telling it to exit if it is still waiting for commands).  Afterwards the instance's tables and
// d_evals are exactly what the host-side bookkeeping (cur, n, evals_valid) says, and regular launches can continue.
int tail_stop(zolt_sc* sc);
}  // namespace
static int sc_bind_launch(zolt_sc* sc, const Fr& r);
namespace {
int tail_stop(zolt_sc* sc) {
  if (!sc->tail_active) return ZOLT_OK;
  if (sc->n >= 2) tail_post(sc, TAIL_OP_EXIT, nullptr, 0);  // after the last bind the kernel has left its loop by itself
  ZOLT_CUDA_TRY(cudaStreamSynchronize(sc->tail_stream));
  sc->tail_active = false;
  sc->tail_pending = false;
  if (sc->tail_bind_seq && *mb_word(sc, kMbCmdAck) < sc->tail_bind_seq) {
    // the kernel retired (it ran out of patience) before it took the last bind: nothing of that command was executed --
    // take it back and run it as a regular launch
    sc->tail_bind_seq = 0;
    sc->cur = sc->tail_prev_cur;
    sc->n = sc->tail_prev_n;
    sc->evals_valid = true;
    return sc_bind_launch(sc, sc->last_r);
  }
  sc->tail_bind_seq = 0;
  return ZOLT_OK;
}

}  // namespace

extern "C" int zolt_sc_create(zolt_ctx* ctx, int combiner, int pairing, int log_n, int n_tables, int n_extra,
                              const uint64_t* const* tables, zolt_sc** out) {
  if (!tables) return ZOLT_ERR_ARG;
  zolt_sc* sc = nullptr;
  ZOLT_TRY(sc_alloc(ctx, combiner, pairing, log_n, n_tables, n_extra, &sc));
  OnFail undo([&] { zolt_sc_free(sc); });
  sc->owns_in = true;
  for (int i = 0; i < n_tables + n_extra; i++) {
    if (!tables[i]) return ZOLT_ERR_ARG;
    Fr* d = nullptr;
    ZOLT_CUDA_TRY(pool_alloc(ctx, sc->in_bytes, (void**)&d));
    sc->tab_in[i] = d;
    ZOLT_CUDA_TRY(cudaMemcpyAsync(d, tables[i], sc->n * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
  }
  ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // host buffers are only borrowed for the call
  undo.dismiss();
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
  if (n_tau < 1 || (sc->combiner == ZOLT_SC_STAGE4_GRUEN ? n_tau > sc->log_n0 : n_tau != sc->log_n0)) return ZOLT_ERR_ARG;
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
  ZOLT_TRY(launch_round_any(sc, false, Fr::zero()));
  ZOLT_CUDA_TRY(cudaGetLastError());
  sc->evals_valid = true;
  return ZOLT_OK;
}

// The round kernel's last thread writes the evaluations into the mapped host mailbox and then the launch's
// sequence number next to them.  Spinning on that word hands the result to the transcript a kernel-retirement
// and a driver round trip earlier than cudaStreamSynchronize (per round, 24 rounds per sumcheck); the stream is
// still queried now and then so that a failed launch or a peer time-out ends the wait.
static int sc_bind_launch(zolt_sc* sc, const Fr& r);
static int sc_wait_evals(zolt_sc* sc) {
  if (!sc->wait_seq) {
    ZOLT_CUDA_TRY(cudaStreamSynchronize(sc->ctx->stream));
    return peer_check(sc->box);
  }
  volatile unsigned* flag = reinterpret_cast<volatile unsigned*>(sc->h_evals + 4);
  const cudaStream_t stream = sc->tail_active ? sc->tail_stream : sc->ctx->stream;
  // cudaStreamQuery is a driver call of several microseconds that delays the very flag it guards: look at the stream
  // only every ~100 us of spinning (r2 v10: queried every 1024 spins, the host noticed a published round 30 us late)
  auto t_last = std::chrono::steady_clock::now();
  for (unsigned spins = 1;; spins++) {
    if (*flag == sc->wait_seq) break;
    if ((spins & 0xffu) == 0 && std::chrono::steady_clock::now() - t_last > std::chrono::microseconds(100)) {
      t_last = std::chrono::steady_clock::now();
      const cudaError_t q = cudaStreamQuery(stream);
      if (q == cudaSuccess) {
        if (*flag == sc->wait_seq) break;
        if (sc->tail_active && sc->tail_pending && *mb_word(sc, kMbTailStatus) == TAIL_ST_TIMEOUT) {
          // the tail kernel gave up waiting before this command was posted (a host that paused for seconds between
          // two rounds): nothing of the command was executed -- take it back and run it as a regular launch
          sc->tail_active = false;
          sc->tail_pending = false;
          sc->tail_bind_seq = 0;
          sc->cur = sc->tail_prev_cur;
          sc->n = sc->tail_prev_n;
          sc->evals_valid = true;
          ZOLT_TRY(sc_bind_launch(sc, sc->last_r));
          return sc_wait_evals(sc);
        }
        ZOLT_TRY(peer_check(sc->box));
        fprintf(stderr, "[zolt_gpu] round kernel retired without delivering its evaluations\n");
        return ZOLT_ERR_CUDA;
      }
      if (q != cudaErrorNotReady) ZOLT_CUDA_TRY(q);
    }
  }
  __atomic_thread_fence(__ATOMIC_ACQUIRE);
  sc->tail_pending = false;
  return ZOLT_OK;
}

extern "C" int zolt_sc_set_claim(zolt_sc* sc, const uint64_t claim[4]) {
  if (!sc || !claim) return ZOLT_ERR_ARG;
  if (sc->cur >= 0 || sc->evals_valid) return ZOLT_ERR_STATE;  // only before the first round is evaluated
  ZOLT_CUDA_TRY(cudaSetDevice(sc->ctx->device));
  ZOLT_CUDA_TRY(cudaMemcpyAsync(sc->d_claim, claim, sizeof(Fr), cudaMemcpyHostToDevice, sc->ctx->stream));
  ZOLT_CUDA_TRY(cudaStreamSynchronize(sc->ctx->stream));  // claim is a borrowed host buffer
  sc->have_claim = true;
  return ZOLT_OK;
}

// The shard protocols (blocked hypercube ranges) are only correct for adjacent pairs and dense combiners: a
// half-split bind pairs an element with one in another rank's block, and the Gruen weights are indexed by the LOCAL pair
// index with tau covering the local variables only.  Refuse instead of returning plausible but wrong polynomials.
static bool sc_shardable(const zolt_sc* sc) {
  return sc->pairing == ZOLT_PAIR_ADJACENT && !combiner_gruen(sc->combiner) && !combiner_stage4(sc->combiner);
}

extern "C" int zolt_sc_set_sharded(zolt_sc* sc, int on) {
  if (!sc) return ZOLT_ERR_ARG;
  if (on && !sc_shardable(sc)) return ZOLT_ERR_ARG;
  ZOLT_TRY(tail_stop(sc));
  if (sc->evals_valid && !sc->evals_final) return ZOLT_ERR_STATE;
  sc->sharded = on != 0;
  return ZOLT_OK;
}

extern "C" int zolt_sc_round_evals(zolt_sc* sc, uint64_t* evals_out) {
  if (!sc || !evals_out) return ZOLT_ERR_ARG;
  if (sc->sharded && !(sc->evals_valid && sc->evals_final)) return ZOLT_ERR_STATE;  // use the partials protocol
  ZOLT_TRY(sc_ensure_evals(sc));
  ZOLT_TRY(sc_wait_evals(sc));
  memcpy(evals_out, sc->h_evals, (size_t)sc->ne * sizeof(Fr));
  return ZOLT_OK;
}

extern "C" int zolt_sc_set_peers(zolt_sc* sc, zolt_peer_box* box) {
  if (!sc || !box || !box->connected || box->ctx != sc->ctx) return ZOLT_ERR_ARG;
  if (!sc_shardable(sc)) return ZOLT_ERR_ARG;
  if (sc->evals_valid || sc->cur >= 0 || sc->sharded) return ZOLT_ERR_STATE;  // before the first round
  sc->box = box;
  return ZOLT_OK;
}

extern "C" int zolt_sc_gather_tail(zolt_sc* sc, const void** d_tables_out, size_t* n_out) {
  if (!sc || !d_tables_out || !n_out || !sc->box) return ZOLT_ERR_ARG;
  ZOLT_TRY(tail_stop(sc));
  if (sc->n < 1 || sc->n > kPeerTailEntries) return ZOLT_ERR_STATE;  // bind further first
  const Fr* ptrs[kMaxTables];
  const Fr* outs[kMaxTables];
  const int total = sc->k + sc->n_extra;
  for (int i = 0; i < total; i++) ptrs[i] = sc->cur < 0 ? sc->tab_in[i] : sc->buf[sc->cur][i];
  ZOLT_TRY(peer_gather_tables(sc->box, ptrs, total, sc->n, outs));
  ZOLT_CUDA_TRY(cudaStreamSynchronize(sc->ctx->stream));
  ZOLT_TRY(peer_check(sc->box));
  for (int i = 0; i < total; i++) d_tables_out[i] = outs[i];
  *n_out = sc->n * (size_t)sc->box->world;
  return ZOLT_OK;
}

extern "C" int zolt_sc_round_partials_dev(zolt_sc* sc, const void** d_ptr) {
  if (!sc || !d_ptr) return ZOLT_ERR_ARG;
  if (!sc->sharded) return ZOLT_ERR_STATE;
  ZOLT_TRY(sc_ensure_evals(sc));
  *d_ptr = sc->d_raw;
  return ZOLT_OK;
}

extern "C" int zolt_sc_finish_partials(zolt_sc* sc, const void* d_rows, size_t rows, uint64_t* evals_out) {
  if (!sc || !d_rows || !evals_out || rows == 0) return ZOLT_ERR_ARG;
  if (!sc->sharded || !sc->evals_valid) return ZOLT_ERR_STATE;
  zolt_ctx* ctx = sc->ctx;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  if (!sc->evals_final) {
    ScOut out{};
    out.d_evals = sc->d_evals;
    out.h_evals = sc->h_evals_dev;
    out.d_claim = sc->d_claim;
    out.d_claim_cur = sc->d_claim + 1;
    out.deg = combiner_deg(sc->combiner);
    out.claim_src = sc->pending_claim_src;
    sc_finish_rows_kernel<<<1, 32, 0, ctx->stream>>>((const Fr*)d_rows, rows, out, sc->last_r);
    ctx->launches++;
    ZOLT_CUDA_TRY(cudaGetLastError());
    sc->evals_final = true;
  }
  ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  memcpy(evals_out, sc->h_evals, (size_t)sc->ne * sizeof(Fr));
  return ZOLT_OK;
}

// one bind through a regular launch (fused bind + next evaluation, or the final bind)
static int sc_bind_launch(zolt_sc* sc, const Fr& r) {
  zolt_ctx* ctx = sc->ctx;
  if (sc->n >= 4) {
    // the fused pass derives p(1) of the next round from this round's polynomial at r
    if (!sc->evals_valid) {
      if (sc->sharded) return ZOLT_ERR_STATE;
      ZOLT_TRY(sc_ensure_evals(sc));
    }
    if (!sc->evals_final) return ZOLT_ERR_STATE;  // sharded: zolt_sc_finish_partials has not run
    sc->last_r = r;
    ZOLT_TRY(launch_round_any(sc, true, r));
    sc->evals_valid = true;
  } else {
    // last variable: bind only, nothing left to evaluate (both pairings pair elements 0 and 1)
    ScPtrs t{};
    const int total = sc->k + sc->n_extra;
    for (int i = 0; i < total; i++) {
      t.src[i] = sc->cur < 0 ? sc->tab_in[i] : sc->buf[sc->cur][i];
      t.dst[i] = sc->buf[sc->cur < 0 ? 0 : (sc->cur ^ 1)][i];
    }
    bind_last_kernel<<<1, 32, 0, ctx->stream>>>(t, total, r);
    ctx->launches++;
    sc->evals_valid = false;
  }
  ZOLT_CUDA_TRY(cudaGetLastError());
  sc->cur = sc->cur < 0 ? 0 : (sc->cur ^ 1);
  sc->n /= 2;
  return ZOLT_OK;
}

extern "C" int zolt_sc_bind(zolt_sc* sc, const uint64_t challenge[4]) {
  if (!sc || !challenge) return ZOLT_ERR_ARG;
  if (sc->n < 2) return ZOLT_ERR_STATE;
  if (combiner_gruen(sc->combiner) && !sc->d_e_out) return ZOLT_ERR_STATE;
  zolt_ctx* ctx = sc->ctx;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  const Fr r = fr_from_host(challenge);
  // small rounds: post the challenge to the resident tail kernel instead of launching (see sc_tail_kernel)
  if (tail_eligible(sc) && sc->n / 4 <= (size_t)ctx->sc_tail_pairs) {
    if (!sc->tail_active && !sc->evals_valid) ZOLT_TRY(sc_ensure_evals(sc));  // the kernel starts from the current round's evaluations
    Fr claim = Fr::zero();
    if (sc->n >= 4 && combiner_deg(sc->combiner) != 0) {
      // the next round's claim is this round's polynomial at r: O(1) host glue on the evaluations in the result mailbox
      // (waited for here if the caller has not fetched them; this also notices a tail kernel that timed out)
      ZOLT_TRY(sc_wait_evals(sc));
      claim = host_cubic_at(sc->h_evals, r);
    } else if (sc->tail_pending) {
      ZOLT_TRY(sc_wait_evals(sc));
    }
    if (!sc->tail_active) {
      const int st = tail_launch(sc);
      if (st == ZOLT_ERR_STATE) return sc_bind_launch(sc, r);  // launches are serialised: one launch per round
      ZOLT_TRY(st);
    }
    sc->last_r = r;
    sc->tail_prev_cur = sc->cur;
    sc->tail_prev_n = sc->n;
    unsigned seq = 0;
    if (sc->n >= 4) {
      if (++ctx->mail_seq == 0) ++ctx->mail_seq;
      seq = sc->wait_seq = ctx->mail_seq;
    }
    tail_post(sc, TAIL_OP_BIND, &r, seq, &claim);
    sc->tail_pending = sc->n >= 4;
    sc->evals_valid = sc->n >= 4;
    sc->evals_final = true;
    sc->cur = sc->cur < 0 ? 0 : (sc->cur ^ 1);
    sc->n /= 2;
    return ZOLT_OK;
  }
  ZOLT_TRY(tail_stop(sc));
  return sc_bind_launch(sc, r);
}

extern "C" int zolt_sc_len(const zolt_sc* sc, size_t* n) {
  if (!sc || !n) return ZOLT_ERR_ARG;
  *n = sc->n;
  return ZOLT_OK;
}

extern "C" int zolt_sc_download(zolt_sc* sc, int table, uint64_t* dst) {
  if (!sc || !dst || table < 0 || table >= sc->k + sc->n_extra) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(sc->ctx->device));
  ZOLT_TRY(tail_stop(sc));  // the resident tail kernel owns the tables until it is joined
  const Fr* src = sc->cur < 0 ? sc->tab_in[table] : sc->buf[sc->cur][table];
  ZOLT_CUDA_TRY(cudaMemcpyAsync(dst, src, sc->n * sizeof(Fr), cudaMemcpyDeviceToHost, sc->ctx->stream));
  ZOLT_CUDA_TRY(cudaStreamSynchronize(sc->ctx->stream));
  return ZOLT_OK;
}

extern "C" int zolt_sc_table_dev(zolt_sc* sc, int table, const void** d_ptr) {
  if (!sc || !d_ptr || table < 0 || table >= sc->k + sc->n_extra) return ZOLT_ERR_ARG;
  ZOLT_TRY(tail_stop(sc));
  *d_ptr = sc->cur < 0 ? sc->tab_in[table] : sc->buf[sc->cur][table];
  return ZOLT_OK;
}

extern "C" int zolt_sc_free(zolt_sc* sc) {
  if (!sc) return ZOLT_ERR_ARG;
  cudaSetDevice(sc->ctx->device);
  tail_stop(sc);
  cudaStreamSynchronize(sc->ctx->stream);
  zolt_ctx* ctx = sc->ctx;
  if (sc->tail_stream) ctx->tail_streams_free.push_back(sc->tail_stream);
  if (sc->tail_ev) cudaEventDestroy(sc->tail_ev);
  for (int i = 0; i < kMaxTables; i++) {
    if (sc->owns_in && sc->tab_in[i]) pool_release(ctx, const_cast<Fr*>(sc->tab_in[i]), sc->in_bytes);
    if (sc->buf[0][i]) pool_release(ctx, sc->buf[0][i], sc->buf_bytes[0]);
    if (sc->buf[1][i]) pool_release(ctx, sc->buf[1][i], sc->buf_bytes[1]);
  }
  if (sc->d_partials) pool_release(ctx, sc->d_partials, (size_t)sc->max_grid * 4 * sizeof(Fr));
  if (sc->d_ticket) pool_release(ctx, sc->d_ticket, 256);
  if (sc->d_evals) pool_release(ctx, sc->d_evals, 4 * sizeof(Fr));
  if (sc->d_raw) pool_release(ctx, sc->d_raw, 4 * sizeof(Fr));
  if (sc->d_claim) pool_release(ctx, sc->d_claim, 4 * sizeof(Fr));
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
    // The scratch comes from the context's buffer cache and goes back without a synchronisation: whoever gets it next
    // works on the same stream, behind the copy.
    Fr* tmp = nullptr;
    ZOLT_CUDA_TRY(pool_alloc(ctx, (n / 2) * sizeof(Fr), (void**)&tmp));
    OnFail give_back([&] { pool_release(ctx, tmp, (n / 2) * sizeof(Fr)); });
    bind_kernel<ZOLT_PAIR_ADJACENT><<<grid, kScThreads, 0, ctx->stream>>>(tmp, (const Fr*)d_src, n, rr);
    ctx->launches++;
    ZOLT_CUDA_TRY(cudaGetLastError());
    ZOLT_CUDA_TRY(cudaMemcpyAsync(d_dst, tmp, (n / 2) * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
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

extern "C" int zolt_eq_plus_one_evals(zolt_ctx* ctx, const uint64_t* r, int n_vars, void* d_out) {
  if (!ctx || !d_out || n_vars < 1 || n_vars > 40 || !r) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  Fr* d_r = nullptr;
  ZOLT_CUDA_TRY(pool_alloc(ctx, 40 * sizeof(Fr), (void**)&d_r));
  OnFail give_back([&] { pool_release(ctx, d_r, 40 * sizeof(Fr)); });
  ZOLT_CUDA_TRY(cudaMemcpyAsync(d_r, r, (size_t)n_vars * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
  eq_plus_one_kernel<<<grid_for(ctx, size_t(1) << n_vars, 8), kScThreads, 0, ctx->stream>>>((Fr*)d_out, d_r, n_vars);
  ctx->launches++;
  ZOLT_CUDA_TRY(cudaGetLastError());
  ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // r is a borrowed host buffer
  return ZOLT_OK;
}

extern "C" int zolt_fr_fold_high_dev(zolt_ctx* ctx, void* d_out, const void* d_in, const void* d_w, size_t n_lo, size_t n_hi,
                                     const uint64_t* scale) {
  if (!ctx || !d_out || !d_in || !d_w || n_lo == 0 || n_hi == 0) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  const size_t tiles = (n_lo + 31) / 32;
  // enough row slices to fill the device twice over, never more slices than groups of rows
  size_t slices = (2 * (size_t)ctx->num_sms + tiles - 1) / tiles;
  const size_t row_groups = (n_hi + kFoldWarps - 1) / kFoldWarps;
  if (slices > row_groups) slices = row_groups;
  if (slices > 1024) slices = 1024;
  if (slices < 1) slices = 1;
  const size_t bytes = (slices * n_lo + 1) * sizeof(Fr);
  Fr* d_part = nullptr;
  ZOLT_CUDA_TRY(pool_alloc(ctx, bytes, (void**)&d_part));
  OnFail give_back([&] { pool_release(ctx, d_part, bytes); });  // runs on every exit
  Fr* d_scale = nullptr;
  if (scale) {
    d_scale = d_part + slices * n_lo;
    ZOLT_CUDA_TRY(cudaMemcpyAsync(d_scale, scale, sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
  }
  fold_high_kernel<<<dim3((unsigned)tiles, (unsigned)slices), 32 * kFoldWarps, 0, ctx->stream>>>((const Fr*)d_in, (const Fr*)d_w, n_lo,
                                                                                             n_hi, d_part);
  fold_finish_kernel<<<(unsigned)((n_lo + kScThreads - 1) / kScThreads), kScThreads, 0, ctx->stream>>>(d_part, (int)slices, n_lo,
                                                                                                    (Fr*)d_out, d_scale);
  ctx->launches += 2;
  ZOLT_CUDA_TRY(cudaGetLastError());
  ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // scale is a borrowed host buffer; the partials go back to the cache
  return ZOLT_OK;
}

extern "C" int zolt_fr_vec_dev(zolt_ctx* ctx, int op, void* d_out, const void* d_a, const void* d_b, const void* d_c, const void* d_e,
                               const uint64_t* s, size_t n) {
  if (!ctx || !d_out) return ZOLT_ERR_ARG;
  if (op == ZOLT_VEC_AFFINE_INDEX) {
    if (!s) return ZOLT_ERR_ARG;
    if (n == 0) return ZOLT_OK;
    ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
    Fr* d_s = nullptr;
    ZOLT_CUDA_TRY(pool_alloc(ctx, 2 * sizeof(Fr), (void**)&d_s));
    ZOLT_CUDA_TRY(cudaMemcpyAsync(d_s, s, 2 * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    vec_kernel<ZOLT_VEC_AFFINE_INDEX><<<grid_for(ctx, n, 8), kScThreads, 0, ctx->stream>>>((Fr*)d_out, d_s, nullptr, nullptr, nullptr,
                                                                                         FixedMulTable{}, n);
    ctx->launches++;
    ZOLT_CUDA_TRY(cudaGetLastError());
    ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // s is a borrowed host buffer
    pool_release(ctx, d_s, 2 * sizeof(Fr));
    return ZOLT_OK;
  }
  if (op == ZOLT_VEC_LIN2 || op == ZOLT_VEC_STAGE4_TERM) {
    if (!d_a || !d_b || (op == ZOLT_VEC_LIN2 ? !s : (!d_c || !d_e))) return ZOLT_ERR_ARG;
    if (n == 0) return ZOLT_OK;
    ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
    const unsigned grid = grid_for(ctx, n, 8);
    if (op == ZOLT_VEC_STAGE4_TERM) {
      vec_kernel<ZOLT_VEC_STAGE4_TERM><<<grid, kScThreads, 0, ctx->stream>>>((Fr*)d_out, (const Fr*)d_a, (const Fr*)d_b, (const Fr*)d_c,
                                                                            (const Fr*)d_e, FixedMulTable{}, n);
      ctx->launches++;
      ZOLT_CUDA_TRY(cudaGetLastError());
      return ZOLT_OK;
    }
    Fr* d_s = nullptr;
    ZOLT_CUDA_TRY(pool_alloc(ctx, 2 * sizeof(Fr), (void**)&d_s));
    ZOLT_CUDA_TRY(cudaMemcpyAsync(d_s, s, 2 * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    vec_kernel<ZOLT_VEC_LIN2><<<grid, kScThreads, 0, ctx->stream>>>((Fr*)d_out, (const Fr*)d_a, (const Fr*)d_b, d_s, nullptr, FixedMulTable{}, n);
    ctx->launches++;
    ZOLT_CUDA_TRY(cudaGetLastError());
    ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // s is a borrowed host buffer
    pool_release(ctx, d_s, 2 * sizeof(Fr));
    return ZOLT_OK;
  }
  if (!d_a) return ZOLT_ERR_ARG;
  if (op == ZOLT_VEC_MUL && !d_b) return ZOLT_ERR_ARG;
  if (op == ZOLT_VEC_EQ_AB_MINUS_C && (!d_b || !d_c || !d_e)) return ZOLT_ERR_ARG;
  if (op == ZOLT_VEC_AXPY && !s) return ZOLT_ERR_ARG;
  if (op < ZOLT_VEC_MUL || op > ZOLT_VEC_AXPY) return ZOLT_ERR_ARG;
  if (n == 0) return ZOLT_OK;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  const unsigned grid = grid_for(ctx, n, 8);
  Fr* o = (Fr*)d_out;
  const Fr *a = (const Fr*)d_a, *b = (const Fr*)d_b, *c = (const Fr*)d_c, *e = (const Fr*)d_e;
  if (op == ZOLT_VEC_MUL) vec_kernel<ZOLT_VEC_MUL><<<grid, kScThreads, 0, ctx->stream>>>(o, a, b, c, e, FixedMulTable{}, n);
  else if (op == ZOLT_VEC_EQ_AB_MINUS_C) vec_kernel<ZOLT_VEC_EQ_AB_MINUS_C><<<grid, kScThreads, 0, ctx->stream>>>(o, a, b, c, e, FixedMulTable{}, n);
  else vec_kernel<ZOLT_VEC_AXPY><<<grid, kScThreads, 0, ctx->stream>>>(o, a, b, c, e, fixed_mul_table_host(fr_from_host(s)), n);
  ctx->launches++;
  ZOLT_CUDA_TRY(cudaGetLastError());
  return ZOLT_OK;
}

namespace {
__global__ void __launch_bounds__(kScThreads)
index_map_kernel(Fr* __restrict__ dst, const Fr* __restrict__ src, size_t n, size_t d1, size_t m1, size_t s1, size_t d2, size_t m2, size_t s2) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    size_t j = 0;
    if (m1) j += ((i / d1) % m1) * s1;
    if (m2) j += ((i / d2) % m2) * s2;
    st_fr(dst + i, ld_fr(src + j));
  }
}
}  // namespace

extern "C" int zolt_fr_index_map_dev(zolt_ctx* ctx, void* d_dst, const void* d_src, size_t n, size_t d1, size_t m1, size_t s1, size_t d2,
                                     size_t m2, size_t s2) {
  if (!ctx || !d_dst || !d_src || d_dst == d_src || (m1 && !d1) || (m2 && !d2)) return ZOLT_ERR_ARG;
  if (n == 0) return ZOLT_OK;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  index_map_kernel<<<grid_for(ctx, n, 8), kScThreads, 0, ctx->stream>>>((Fr*)d_dst, (const Fr*)d_src, n, d1, m1, s1, d2, m2, s2);
  ctx->launches++;
  ZOLT_CUDA_TRY(cudaGetLastError());
  return ZOLT_OK;
}

extern "C" int zolt_poly_sum(zolt_ctx* ctx, const void* d_src, size_t n, uint64_t out[4]) {
  if (!ctx || !d_src || !out) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  Scratch s;
  ZOLT_TRY(scratch_alloc(ctx, &s));
  ScOut so{};
  so.partials = s.partials;
  so.ticket = s.ticket;
  so.d_raw = s.d_out + 8;
  so.d_evals = s.d_out;
  so.h_evals = (Fr*)ctx->d_mailbox;
  so.finalize = 1;
  sum_kernel<<<grid_for(ctx, n, 4), kScThreads, 0, ctx->stream>>>((const Fr*)d_src, n, so);
  ctx->launches++;
  ZOLT_CUDA_TRY(cudaGetLastError());
  ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  memcpy(out, ctx->h_mailbox, 32);
  scratch_free(ctx, &s);
  return ZOLT_OK;
}

extern "C" int zolt_fr_inner_product_dev(zolt_ctx* ctx, const void* d_a, const void* d_b, size_t n, uint64_t out[4]) {
  if (!ctx || !d_a || !d_b || !out) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  Scratch s;
  ZOLT_TRY(scratch_alloc(ctx, &s));
  ScOut so{};
  so.partials = s.partials;
  so.ticket = s.ticket;
  so.d_raw = s.d_out + 8;
  so.d_evals = s.d_out;
  so.h_evals = (Fr*)ctx->d_mailbox;
  so.finalize = 1;
  inner_product_kernel<<<grid_for(ctx, n, 4), kScThreads, 0, ctx->stream>>>((const Fr*)d_a, (const Fr*)d_b, n, so);
  ctx->launches++;
  ZOLT_CUDA_TRY(cudaGetLastError());
  ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  memcpy(out, ctx->h_mailbox, 32);
  scratch_free(ctx, &s);
  return ZOLT_OK;
}

extern "C" int zolt_lasso_addr_round(zolt_ctx* ctx, void* d_eq_evals, const void* d_indices, size_t n, int bind_bit, const uint64_t* r,
                                     int sum_bit, uint64_t sums_out[8]) {
  if (!ctx || !d_eq_evals || !d_indices || !sums_out || bind_bit > 127 || sum_bit > 127 || (bind_bit >= 0 && !r)) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  Scratch s;
  ZOLT_TRY(scratch_alloc(ctx, &s));
  OnFail give_back([&] { scratch_free(ctx, &s); });  // runs on every exit
  ScOut so{};
  so.partials = s.partials;
  so.ticket = s.ticket;
  so.d_raw = s.d_out + 8;
  so.d_evals = s.d_out;
  so.h_evals = (Fr*)ctx->d_mailbox;
  so.finalize = 1;
  const FixedMulTable rt = bind_bit >= 0 ? fixed_mul_table_host(fr_from_host(r)) : FixedMulTable{};
  lasso_addr_kernel<<<grid_for(ctx, n, 4), kScThreads, 0, ctx->stream>>>((Fr*)d_eq_evals, (const uint4*)d_indices, n, bind_bit < 0 ? -1 : bind_bit,
                                                                        rt, sum_bit < 0 ? -1 : sum_bit, so);
  ctx->launches++;
  ZOLT_CUDA_TRY(cudaGetLastError());
  ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  memcpy(sums_out, ctx->h_mailbox, 64);
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
