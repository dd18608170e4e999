// BN254 G1 multi-scalar multiplication for sm_100a.
//
// Replaces (reference, /root/reference): MSM(F,G).compute / pippengerMSM (src/msm/mod.zig:355-438),
// BatchMSM (:545-565), ParallelMSM's chunk/combine split (:588-653), and behind them
// HyperKZG.commit (src/poly/commitment/mod.zig:239-255), HyperKZG.setup (:174-213) and the Dory
// row commitments (src/poly/commitment/dory.zig:646-670).
//
// Pipeline (k MSMs over the same n bases share every stage; key = (batch, window, bucket)):
//   1. digits   : Montgomery -> canonical scalar, signed c-bit digits, per-bucket histogram
//   2. scan     : exclusive scan of the histogram -> bucket offsets
//   3. scatter  : counting sort of (point index, sign) by key
//   4. accumulate (level 0): each thread folds a fixed-length slice of the SORTED list with mixed
//      additions; groups interior to a slice are complete buckets, the first/last group of a slice
//      may continue in a neighbour and go to an "edge" list
//   5. accumulate (levels 1..): the edge list (still sorted by key) is folded the same way with
//      full additions until one thread sees everything.  Work per thread is bounded by the slice
//      length whatever the digit distribution (all-equal scalars included).
//   6. bucket reduction: chunked running sums per window, tree over chunks, Horner over windows
//   7. normalisation to canonical affine
// Bucket contents are group elements, so the order in which equal-key points are folded does not
// change the canonical affine result.
#include <cstring>
#include <new>

#include <atomic>
#include "common.cuh"
#include "curve.cuh"
#include "peer.cuh"

using namespace zolt;

struct zolt_bases {
  zolt_ctx* ctx = nullptr;
  size_t n = 0;
  AffineG1* d_xy = nullptr;
  uint8_t* d_inf = nullptr;  // nullptr: no point at infinity in the set
  // Optional window-level table (zolt_msm_bases_precompute): d_pre[j*n + i] = 2^(pre_c*j) * base[i], affine,
  // j < pre_w.  With it every window's digit indexes ONE shared bucket set: a W-window MSM over n points
  // becomes a single-window MSM over W*n table points -- no per-window bucket reduction, no Horner
  // doublings, and a window width chosen for the addition count alone (22 bits -> 12 additions per point
  // instead of 16).  12.9 GB at n = 2^24: HBM3e capacity is what buys the 25 % of integer work.
  AffineG1* d_pre = nullptr;
  int pre_c = 0, pre_w = 0;
  bool borrowed = false;  // zolt_msm_bases_wrap_dev: d_xy / d_inf belong to the caller
};

namespace {

constexpr uint32_t kInvalidKey = 0xffffffffu;
constexpr int kSlice0 = 128;     // level-0 slice length (mixed additions per thread)
constexpr int kSlice0Big = 256;  // ... for lists of 2^26 entries and more: half the edge entries, still >= 14 waves of blocks
constexpr int kSliceE = 8;    // edge-level slice length (full additions per thread)
constexpr int kChunk = 16;    // buckets per thread in the running-sum reduction

struct MsmGeom {
  int c;         // window bits
  int windows;   // ceil(255 / c)
  uint32_t B;    // buckets per window = 2^(c-1)
};

// Window size by problem size (per MSM of the batch).  Cost model: windows * n mixed additions +
// windows * B * ~3 addition-equivalents of reduction.
MsmGeom choose_geom(size_t n) {
  int c;
  if (n < (size_t(1) << 8)) c = 5;
  else if (n < (size_t(1) << 10)) c = 7;
  else if (n < (size_t(1) << 12)) c = 8;
  else if (n < (size_t(1) << 14)) c = 10;
  else if (n < (size_t(1) << 16)) c = 11;
  else if (n < (size_t(1) << 18)) c = 13;
  else if (n < (size_t(1) << 20)) c = 14;
  else if (n < (size_t(1) << 22)) c = 15;
  else c = 16;
  MsmGeom g;
  g.c = c;
  g.windows = (255 + c - 1) / c;
  g.B = 1u << (c - 1);
  return g;
}

__device__ __forceinline__ Fq ld_fq(const Fq* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  Fq r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void st_fq(Fq* p, const Fq& x) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
  q[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}
__device__ __forceinline__ AffineG1 ld_affine(const AffineG1* p) {
  AffineG1 r;
  r.x = ld_fq(&p->x);
  r.y = ld_fq(&p->y);
  return r;
}
__device__ __forceinline__ XYZZ ld_xyzz(const XYZZ* p) {
  XYZZ r;
  r.x = ld_fq(&p->x);
  r.y = ld_fq(&p->y);
  r.zz = ld_fq(&p->zz);
  r.zzz = ld_fq(&p->zzz);
  return r;
}
// coherent (L2) load for data produced earlier in the same kernel by other threads
__device__ __forceinline__ XYZZ ld_xyzz_cg(const XYZZ* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  XYZZ r;
  uint32_t* w = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint4 t = __ldcg(q + i);
    w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w;
  }
  return r;
}
__device__ __forceinline__ void st_xyzz(XYZZ* p, const XYZZ& v) {
  st_fq(&p->x, v.x);
  st_fq(&p->y, v.y);
  st_fq(&p->zz, v.zz);
  st_fq(&p->zzz, v.zzz);
}

// ---- 1. digits + histogram ---------------------------------------------------------------------------
// digits[(batch*W + w)*n + i] = |d| | sign << 31 (0: nothing to add).  Signed recoding: d in
// (-2^(c-1), 2^(c-1)], a carry moves to the next window; windows*c >= 255 leaves room for the
// carry of a 254-bit scalar.  getWindow's per-call fromMontgomery (src/msm/mod.zig:441-442) becomes
// one conversion per scalar.
__global__ void __launch_bounds__(256)
msm_digits_kernel(const Fr* __restrict__ scalars, const uint8_t* __restrict__ inf, size_t n, size_t k,
                  MsmGeom g, int pre, uint32_t* __restrict__ digits, uint32_t* __restrict__ counts) {
  const size_t total = n * k;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const size_t batch = t / n, i = t - batch * n;
    Fr s = fp_from_mont<FrParams>(scalars[t]);
    const bool skip = (inf != nullptr && inf[i] != 0);  // identity bases are skipped (msm/mod.zig:407)
    uint32_t carry = 0;
    const uint32_t mask = (1u << g.c) - 1u;
    for (int w = 0; w < g.windows; w++) {
      const int pos = w * g.c;
      const int word = pos >> 5, sh = pos & 31;
      uint64_t two = (uint64_t)s.v[word < 8 ? word : 7];
      if (word >= 8) two = 0;
      if (word + 1 < 8) two |= (uint64_t)s.v[word + 1] << 32;
      uint32_t d = ((uint32_t)(two >> sh) & mask) + carry;
      uint32_t neg = 0;
      if (d > g.B) {
        d = (1u << g.c) - d;
        neg = 1;
        carry = 1;
      } else {
        carry = 0;
      }
      if (skip) d = 0;
      const size_t row = batch * g.windows + w;
      const size_t krow = pre ? batch : row;  // precomputed levels: all windows share one bucket set
      digits[row * n + i] = d ? (d | (neg << 31)) : 0u;
      if (d) atomicAdd(&counts[krow * g.B + (d - 1)], 1u);
    }
  }
}

// ---- 2. exclusive scan (u32), three passes ---------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;  // per thread
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* smem, uint32_t* total) {
  // smem: kScanThreads entries
  const int tid = threadIdx.x;
  smem[tid] = v;
  __syncthreads();
  for (int off = 1; off < kScanThreads; off <<= 1) {
    uint32_t x = tid >= off ? smem[tid - off] : 0;
    __syncthreads();
    smem[tid] += x;
    __syncthreads();
  }
  uint32_t incl = smem[tid];
  if (total) *total = smem[kScanThreads - 1];
  __syncthreads();
  return incl - v;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(const uint32_t* in, size_t n, uint32_t* tile_sums) {
  __shared__ uint32_t smem[kScanThreads];
  const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; j++)
    if (base + j < n) s += in[base + j];
  uint32_t total;
  block_exclusive_scan(s, smem, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: exclusive scan of tile sums in place; writes the grand total to tile_sums[num_tiles]
__global__ void __launch_bounds__(kScanThreads) scan_tile_offsets_kernel(uint32_t* tile_sums, size_t num_tiles) {
  __shared__ uint32_t smem[kScanThreads];
  __shared__ uint32_t running;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  for (size_t base = 0; base < num_tiles; base += kScanThreads) {
    const size_t i = base + threadIdx.x;
    uint32_t v = i < num_tiles ? tile_sums[i] : 0;
    uint32_t total;
    uint32_t ex = block_exclusive_scan(v, smem, &total);
    if (i < num_tiles) tile_sums[i] = running + ex;
    __syncthreads();
    if (threadIdx.x == 0) running += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) tile_sums[num_tiles] = running;
}

__global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(const uint32_t* in, size_t n, const uint32_t* tile_offsets, uint32_t* out /* n+1 */) {
  __shared__ uint32_t smem[kScanThreads];
  const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; j++) {
    v[j] = (base + j < n) ? in[base + j] : 0;
    s += v[j];
  }
  uint32_t ex = block_exclusive_scan(s, smem, nullptr) + tile_offsets[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kScanItems; j++) {
    if (base + j < n) out[base + j] = ex;
    ex += v[j];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = tile_offsets[gridDim.x];
}

// ---- 3. scatter ----------------------------------------------------------------------------------------
// Each thread keeps kScatterE independent (load, atomic, store) chains in flight: the kernel is bound by the
// round trip of the returning atomic, not by bandwidth (ncu r1 v9: long-scoreboard stalls, DRAM at 10 %).
constexpr int kScatterE = 4;
// grid.y = digit row (batch * windows + w): row, batch and w are per-block constants, no 64-bit division per entry
__global__ void __launch_bounds__(256)
msm_scatter_kernel(const uint32_t* __restrict__ digits, size_t n, size_t rows, MsmGeom g, int pre, size_t pre_n,
                   size_t pre_offset, uint32_t* __restrict__ cursor, uint32_t* __restrict__ sorted) {
  for (size_t row = blockIdx.y; row < rows; row += gridDim.y) {
    size_t krow = row, pbase = 0;
    if (pre) {
      const size_t batch = row / g.windows, w = row - batch * g.windows;
      krow = batch;
      pbase = w * pre_n + pre_offset;  // index of point 0 in the level table of window w
    }
    const uint32_t* __restrict__ drow = digits + row * n;
    uint32_t* __restrict__ crow = cursor + krow * g.B;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i0 < n; i0 += stride * kScatterE) {
      uint32_t d[kScatterE], pos[kScatterE];
#pragma unroll
      for (int j = 0; j < kScatterE; j++) {
        const size_t i = i0 + (size_t)j * stride;
        d[j] = i < n ? __ldcs(drow + i) : 0u;
      }
#pragma unroll
      for (int j = 0; j < kScatterE; j++)
        if (d[j]) pos[j] = atomicAdd(&crow[(d[j] & 0x7fffffffu) - 1], 1u);  // cursor starts as a copy of the bucket offsets
#pragma unroll
      for (int j = 0; j < kScatterE; j++)
        if (d[j]) sorted[pos[j]] = (uint32_t)(pbase + i0 + (size_t)j * stride) | (d[j] & 0x80000000u);
    }
  }
}

// ---- 3b. two-pass bucket sort (one shared bucket set: window-level table, k == 1) ---------------------------------
// The scatter above pays one returning global atomic and one 4-byte store to a random address per entry: 2 x 218 M L2
// round trips at 2^24 points (4.5 ms), with a write frontier (one open sector per bucket) far larger than L2.  Here the
// entries are first partitioned into BINS of 512 consecutive buckets, then every bin is sorted in chunks that fit shared
// memory, so that
//   * all per-entry atomics are SHARED-memory atomics (a block owns its cursors: pass 1 reserved its runs with one exclusive
//     scan over the [bin][block] histogram; pass 2 reserves one range per (bucket, chunk) -- 512 global atomics per 16384 entries);
//   * pass 1 sorts a tile of 1024 scalars' entries by bin in shared memory and appends every bin's piece (13 records = 104 bytes
//     on average) to the block's run of that bin;
//   * pass 2 writes every bucket's share of a chunk as one contiguous run (32 entries = 128 bytes on average).
// 2^24 points: 3.4 ms instead of 5.8 ms for digits + scatter (MSM 39.1 -> 36.7 ms); neutral from 2^18 to 2^21 points, faster on
// trace-like scalars at every size (many equal digits: the atomic scatter serialises on a few cursors).
// The digits are never materialised: both partition kernels re-derive them from the scalars (13 shifts per scalar).
// Output is what the accumulation kernel expects: `sorted` (point index | sign) grouped by bucket, and `offsets`.
constexpr int kBinSubBits = 9;                 // buckets per bin (the open 128-byte lines of pass 1, blocks x bins, must fit L2)
constexpr int kBinSub = 1 << kBinSubBits;
constexpr int kBinThreads = 1024;
constexpr int kChunkRecs = 16384;              // records per pass-2 chunk: 32 per bucket on average, a full 128-byte line
constexpr int kChunkThreads = 1024;
constexpr int kChunkPer = kChunkRecs / kChunkThreads;

// exclusive scan over the block (any multiple of 32 threads up to 1024); `sm` holds 33 words; total in sm[32] afterwards
__device__ __forceinline__ uint32_t block_scan_excl(uint32_t v, uint32_t* sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += y;
  }
  if (lane == 31) sm[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const uint32_t t = lane < (int)(blockDim.x >> 5) ? sm[lane] : 0;
    uint32_t ti = t;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, ti, d);
      if (lane >= d) ti += y;
    }
    sm[lane] = ti - t;
    if (lane == 31) sm[32] = ti;
  }
  __syncthreads();
  return sm[warp] + incl - v;
}

// calls f(w, magnitude 1..B, negative) for every non-zero signed digit of the (standard-form) scalar s
template <class F>
__device__ __forceinline__ void for_each_digit(const Fr& s, const MsmGeom& g, F&& f) {
  uint32_t carry = 0;
  const uint32_t mask = (1u << g.c) - 1u;
  for (int w = 0; w < g.windows; w++) {
    const int pos = w * g.c;
    const int word = pos >> 5, sh = pos & 31;
    uint64_t two = (uint64_t)s.v[word < 8 ? word : 7];
    if (word >= 8) two = 0;
    if (word + 1 < 8) two |= (uint64_t)s.v[word + 1] << 32;
    uint32_t d = ((uint32_t)(two >> sh) & mask) + carry;
    uint32_t neg = 0;
    if (d > g.B) {
      d = (1u << g.c) - d;
      neg = 1;
      carry = 1;
    } else {
      carry = 0;
    }
    if (d) f(w, d, neg);
  }
}

// pass 1a: histogram of this block's share of the scalars over the bins -> matrix[bin * gridDim.x + block]
__global__ void __launch_bounds__(kBinThreads)
msm_bin_hist_kernel(const Fr* __restrict__ scalars, const uint8_t* __restrict__ inf, size_t n, MsmGeom g, uint32_t nbins,
                    uint32_t* __restrict__ matrix) {
  extern __shared__ uint32_t bin_smem[];
  for (uint32_t b = threadIdx.x; b < nbins; b += blockDim.x) bin_smem[b] = 0;
  __syncthreads();
  const size_t per = (n + gridDim.x - 1) / gridDim.x;
  const size_t lo = blockIdx.x * per, hi = lo + per < n ? lo + per : n;
  for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    if (inf != nullptr && inf[i] != 0) continue;  // identity bases are skipped (msm/mod.zig:407)
    const Fr s = fp_from_mont<FrParams>(scalars[i]);
    for_each_digit(s, g, [&](int, uint32_t d, uint32_t) { atomicAdd(&bin_smem[(d - 1) >> kBinSubBits], 1u); });
  }
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < nbins; b += blockDim.x) matrix[(size_t)b * gridDim.x + blockIdx.x] = bin_smem[b];
}

// pass 1b: the same walk in tiles of blockDim.x scalars.  A tile's entries are sorted by bin in shared memory and every
// bin's share is appended to this block's run of that bin as ONE contiguous piece (the run cursors start at the exclusive
// scan of the matrix and live in shared memory).  Writing each entry straight to its run -- 218 M scattered 8-byte stores --
// ran at the DRAM's random-access rate, like the atomic scatter (ncu: 4.0 GB written and 1.7 GB read back for 1.75 GB of
// records, 5.3 ms at 2^24): the open lines of (blocks x bins) runs do not survive in L2 between two touches.
constexpr int kBinMaxWindows = 16;
__global__ void __launch_bounds__(kBinThreads)
msm_bin_partition_kernel(const Fr* __restrict__ scalars, const uint8_t* __restrict__ inf, size_t n, MsmGeom g, uint32_t nbins,
                         const uint32_t* __restrict__ matrix_off, size_t pre_n, size_t pre_offset, uint2* __restrict__ recs) {
  // dynamic shared memory: cursor[nbins] | cnt[nbins] | loff[nbins] | stage[blockDim * windows] (uint2) | stage_bin[...] (u16)
  extern __shared__ uint32_t bin_smem[];
  uint32_t* const cursor = bin_smem;
  uint32_t* const cnt = cursor + nbins;
  uint32_t* const loff = cnt + nbins;
  uint2* const stage = reinterpret_cast<uint2*>(loff + nbins);
  uint16_t* const stage_bin = reinterpret_cast<uint16_t*>(stage + (size_t)blockDim.x * g.windows);
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t pass_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (uint32_t b = tid; b < nbins; b += blockDim.x) {
    cursor[b] = matrix_off[(size_t)b * gridDim.x + blockIdx.x];
    cnt[b] = 0;
  }
  __syncthreads();
  const size_t per = (n + gridDim.x - 1) / gridDim.x;
  const size_t lo = blockIdx.x * per, hi = lo + per < n ? lo + per : n;
  Fr next = lo + tid < hi ? scalars[lo + tid] : Fr::zero();
  for (size_t base = lo; base < hi; base += blockDim.x) {
    const size_t i = base + tid;
    const Fr cur = next;
    if (i + blockDim.x < hi) next = scalars[i + blockDim.x];  // the next tile's scalar travels while this tile is sorted
    uint32_t dg[kBinMaxWindows];  // bucket | sign << 31, or 0xffffffff: no entry
    uint16_t rank[kBinMaxWindows];
#pragma unroll
    for (int w = 0; w < kBinMaxWindows; w++) dg[w] = 0xffffffffu;
    if (i < hi && !(inf != nullptr && inf[i] != 0)) {
      const Fr s = fp_from_mont<FrParams>(cur);
      uint32_t carry = 0;
      const uint32_t mask = (1u << g.c) - 1u;
#pragma unroll
      for (int w = 0; w < kBinMaxWindows; w++) {
        if (w < g.windows) {
          const int pos = w * g.c;
          const int word = pos >> 5, sh = pos & 31;
          uint64_t two = word < 8 ? (uint64_t)s.v[word < 8 ? word : 7] : 0;
          if (word + 1 < 8) two |= (uint64_t)s.v[word + 1 < 8 ? word + 1 : 7] << 32;
          uint32_t d = ((uint32_t)(two >> sh) & mask) + carry;
          uint32_t neg = 0;
          if (d > g.B) {
            d = (1u << g.c) - d;
            neg = 1;
            carry = 1;
          } else {
            carry = 0;
          }
          if (d) dg[w] = (d - 1) | (neg << 31);
        }
      }
#pragma unroll
      for (int w = 0; w < kBinMaxWindows; w++)
        if (dg[w] != 0xffffffffu) rank[w] = (uint16_t)atomicAdd(&cnt[(dg[w] & 0x7fffffffu) >> kBinSubBits], 1u);
    }
    __syncthreads();
    // exclusive scan of cnt[0..nbins) into loff, blockDim.x counters per pass
    uint32_t running = 0;
    for (uint32_t b0 = 0; b0 < nbins; b0 += blockDim.x) {
      const uint32_t b = b0 + tid;
      const uint32_t c = b < nbins ? cnt[b] : 0;
      uint32_t incl = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += y;
      }
      if (lane == 31) warp_tot[warp] = incl;
      __syncthreads();
      if (warp == 0) {
        const uint32_t t = lane < (int)(blockDim.x >> 5) ? warp_tot[lane] : 0;
        uint32_t ti = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t y = __shfl_up_sync(0xffffffffu, ti, d);
          if (lane >= d) ti += y;
        }
        warp_tot[lane] = ti - t;  // exclusive prefix of every warp
        if (lane == 31) pass_total = ti;
      }
      __syncthreads();
      if (b < nbins) loff[b] = running + warp_tot[warp] + incl - c;
      running += pass_total;
      __syncthreads();  // warp_tot / pass_total are rewritten by the next pass
    }
    // stage the entries in bin order
#pragma unroll
    for (int w = 0; w < kBinMaxWindows; w++)
      if (dg[w] != 0xffffffffu) {
        const uint32_t bin = (dg[w] & 0x7fffffffu) >> kBinSubBits;
        const uint32_t slot = loff[bin] + rank[w];
        stage[slot] = make_uint2((uint32_t)((size_t)w * pre_n + pre_offset + i) | (dg[w] & 0x80000000u), dg[w] & 0x7fffffffu);
        stage_bin[slot] = (uint16_t)bin;
      }
    __syncthreads();
    // every bin's piece goes out contiguously behind what this block has already written for that bin
    for (uint32_t j = tid; j < running; j += blockDim.x) {
      const uint32_t bin = stage_bin[j];
      recs[cursor[bin] + (j - loff[bin])] = stage[j];
    }
    __syncthreads();
    for (uint32_t b = tid; b < nbins; b += blockDim.x) {
      cursor[b] += cnt[b];
      cnt[b] = 0;
    }
    __syncthreads();
  }
}

// chunk_base[b] = number of pass-2 chunks before bin b (bin b has ceil(size / kChunkRecs) of them); one block
__global__ void __launch_bounds__(kScanThreads)
msm_chunk_index_kernel(const uint32_t* __restrict__ matrix_off, uint32_t nbins, uint32_t G, uint32_t* __restrict__ chunk_base) {
  __shared__ uint32_t smem[kScanThreads];
  const uint32_t per = (nbins + kScanThreads - 1) / kScanThreads;
  const uint32_t b0 = threadIdx.x * per, b1 = b0 + per < nbins ? b0 + per : nbins;
  uint32_t s = 0;
  for (uint32_t b = b0; b < b1; b++) {
    const uint32_t sz = matrix_off[(size_t)(b + 1) * G] - matrix_off[(size_t)b * G];  // matrix_off has nbins * G + 1 entries
    s += (sz + kChunkRecs - 1) / kChunkRecs;
  }
  uint32_t total;
  uint32_t ex = block_exclusive_scan(s, smem, &total);
  for (uint32_t b = b0; b < b1; b++) {
    chunk_base[b] = ex;
    const uint32_t sz = matrix_off[(size_t)(b + 1) * G] - matrix_off[(size_t)b * G];
    ex += (sz + kChunkRecs - 1) / kChunkRecs;
  }
  if (threadIdx.x == 0) chunk_base[nbins] = total;
}

// (bin, record range) of chunk `chunk`; false when there is no such chunk
__device__ __forceinline__ bool chunk_range(const uint32_t* __restrict__ chunk_base, const uint32_t* __restrict__ matrix_off,
                                            uint32_t nbins, uint32_t G, uint32_t chunk, uint32_t* bin, uint32_t* start, uint32_t* end) {
  if (chunk >= chunk_base[nbins]) return false;
  uint32_t lo = 0, hi = nbins;  // invariant: chunk_base[lo] <= chunk < chunk_base[hi]
  while (hi - lo > 1) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (chunk_base[mid] <= chunk) lo = mid; else hi = mid;
  }
  const uint32_t b0 = matrix_off[(size_t)lo * G], b1 = matrix_off[(size_t)(lo + 1) * G];
  *bin = lo;
  *start = b0 + (chunk - chunk_base[lo]) * kChunkRecs;
  *end = *start + kChunkRecs < b1 ? *start + kChunkRecs : b1;
  return true;
}

// pass 2a: bucket histogram, one shared-memory count per chunk folded into `counts` with at most 128 global adds
__global__ void __launch_bounds__(kChunkThreads)
msm_chunk_count_kernel(const uint2* __restrict__ recs, const uint32_t* __restrict__ chunk_base, const uint32_t* __restrict__ matrix_off,
                       uint32_t nbins, uint32_t G, uint32_t* __restrict__ counts) {
  __shared__ uint32_t cnt[kBinSub];
  __shared__ uint32_t rng[3];
  if (threadIdx.x == 0) {
    uint32_t bin = 0, start = 0, end = 0;
    if (!chunk_range(chunk_base, matrix_off, nbins, G, blockIdx.x, &bin, &start, &end)) end = start = 0;
    rng[0] = bin; rng[1] = start; rng[2] = end;
  }
  for (int q = threadIdx.x; q < kBinSub; q += kChunkThreads) cnt[q] = 0;
  __syncthreads();
  const uint32_t bin = rng[0], start = rng[1], end = rng[2];
  if (start == end) return;
  for (uint32_t p = start + threadIdx.x; p < end; p += kChunkThreads) atomicAdd(&cnt[__ldcs(&recs[p]).y & (kBinSub - 1)], 1u);
  __syncthreads();
  for (int q = threadIdx.x; q < kBinSub; q += kChunkThreads)
    if (cnt[q]) atomicAdd(&counts[bin * kBinSub + q], cnt[q]);
}

// pass 2b: the chunk sorted by bucket in shared memory, every bucket's run copied to the range it reserves in `sorted`
__global__ void __launch_bounds__(kChunkThreads)
msm_chunk_place_kernel(const uint2* __restrict__ recs, const uint32_t* __restrict__ chunk_base, const uint32_t* __restrict__ matrix_off,
                       uint32_t nbins, uint32_t G, uint32_t* __restrict__ cursor, uint32_t* __restrict__ sorted) {
  __shared__ uint32_t cnt[kBinSub], loff[kBinSub + 1], gpos[kBinSub];
  extern __shared__ uint32_t place_smem[];  // kChunkRecs values, then kChunkRecs 16-bit bucket numbers (dynamic: > 48 KB)
  uint32_t* const stage = place_smem;
  uint16_t* const stage_b = reinterpret_cast<uint16_t*>(place_smem + kChunkRecs);
  __shared__ uint32_t rng[3];
  __shared__ uint32_t scan_sm[33];
  const int tid = threadIdx.x;
  if (tid == 0) {
    uint32_t bin = 0, start = 0, end = 0;
    if (!chunk_range(chunk_base, matrix_off, nbins, G, blockIdx.x, &bin, &start, &end)) end = start = 0;
    rng[0] = bin; rng[1] = start; rng[2] = end;
  }
  for (int q = tid; q < kBinSub; q += kChunkThreads) cnt[q] = 0;
  __syncthreads();
  const uint32_t bin = rng[0], start = rng[1], end = rng[2];
  if (start == end) return;
  uint32_t val[kChunkPer];
  uint16_t rank[kChunkPer];
  uint16_t bk[kChunkPer];
#pragma unroll
  for (int j = 0; j < kChunkPer; j++) {  // all loads of the thread in flight together ...
    const uint32_t p = start + j * kChunkThreads + tid;
    uint2 r = make_uint2(0u, 0u);
    if (p < end) r = __ldcs(&recs[p]);
    val[j] = r.x;
    bk[j] = (uint16_t)(r.y & (kBinSub - 1));
  }
#pragma unroll
  for (int j = 0; j < kChunkPer; j++) {  // ... then the ranks
    const uint32_t p = start + j * kChunkThreads + tid;
    if (p < end) rank[j] = (uint16_t)atomicAdd(&cnt[bk[j]], 1u);
  }
  __syncthreads();
  {
    // exclusive scan of the kBinSub counts (one per thread, the rest of the block scans zeros), then one global
    // reservation per non-empty bucket
    static_assert(kBinSub <= kChunkThreads, "one count per thread");
    const uint32_t c = tid < kBinSub ? cnt[tid] : 0;
    const uint32_t ex = block_scan_excl(c, scan_sm);
    if (tid < kBinSub) {
      loff[tid] = ex;
      gpos[tid] = c ? atomicAdd(&cursor[bin * kBinSub + tid], c) : 0u;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kChunkPer; j++) {
    const uint32_t p = start + j * kChunkThreads + tid;
    if (p < end) {
      const uint32_t slot = loff[bk[j]] + rank[j];
      stage[slot] = val[j];
      stage_b[slot] = bk[j];
    }
  }
  __syncthreads();
  const uint32_t count = end - start;
  for (uint32_t j = tid; j < count; j += kChunkThreads) {
    const uint32_t b = stage_b[j];
    sorted[gpos[b] + (j - loff[b])] = stage[j];
  }
}

// ---- 4. accumulate, level 0 ---------------------------------------------------------------------------
// Thread t folds sorted[t*S, (t+1)*S).  Keys are implicit: entry p belongs to the bucket k with
// offsets[k] <= p < offsets[k+1].
__global__ void __launch_bounds__(128, 4)
msm_accum_points_kernel(const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ offsets, uint32_t num_keys,
                        const AffineG1* __restrict__ bases, XYZZ* __restrict__ buckets,
                        uint32_t* __restrict__ edge_keys, XYZZ* __restrict__ edge_pts, size_t num_slices, uint32_t slice_len,
                        unsigned long long* __restrict__ dbg) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (dbg && (threadIdx.x & 31) == 0) {  // diagnostics (tuning key 8): per warp start, end, SM
    unsigned long long tt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
    dbg[3 * (t >> 5)] = tt;
  }
  if (t >= num_slices) return;
  const uint32_t total = offsets[num_keys];
  const uint64_t p0 = (uint64_t)t * slice_len;
  uint32_t ek0 = kInvalidKey, ek1 = kInvalidKey;
  if (p0 < total) {
    const uint32_t p1 = (uint32_t)min((uint64_t)total, p0 + slice_len);
    // largest k with offsets[k] <= p0 (skips empty buckets automatically)
    uint32_t lo = 0, hi = num_keys;  // invariant: offsets[lo] <= p0 < offsets[hi]
    while (hi - lo > 1) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if (offsets[mid] <= (uint32_t)p0) lo = mid; else hi = mid;
    }
    uint32_t key = lo;
    uint32_t boundary = offsets[key + 1];  // first position that belongs to a later bucket
    bool first = true;
    XYZZ acc = XYZZ::identity();
    // One flat loop over the slice: every lane performs exactly one mixed addition per iteration, so
    // the warp stays converged although its lanes cross bucket boundaries at different positions.
    // The point of iteration i+1 is fetched while iteration i computes.
    uint32_t e = sorted[(uint32_t)p0];
    AffineG1 pt = ld_affine(bases + (e & 0x7fffffffu));
    for (uint32_t pos = (uint32_t)p0; pos < p1; pos++) {
      const uint32_t e_cur = e;
      AffineG1 cur = pt;
      if (pos + 1 < p1) {
        e = sorted[pos + 1];
        pt = ld_affine(bases + (e & 0x7fffffffu));
      }
      if (pos == boundary) {  // bucket `key` ended at pos: publish it and move on
        if (first) {
          edge_keys[2 * t] = key;  // may continue in the previous slice
          st_xyzz(&edge_pts[2 * t], acc);
          ek0 = key;
          first = false;
        } else {
          st_xyzz(&buckets[key], acc);  // strictly inside this slice: complete
        }
        acc = XYZZ::identity();
        key++;
        while (offsets[key + 1] == pos) key++;  // skip empty buckets
        boundary = offsets[key + 1];
      }
      if (e_cur >> 31) cur.y = fp_neg<FqParams>(cur.y);
      xyzz_madd_uniform(acc, cur);
    }
    // the bucket open at the end of the slice may continue in the next slice
    if (first) {
      edge_keys[2 * t] = key;
      st_xyzz(&edge_pts[2 * t], acc);
      ek0 = key;
    } else {
      edge_keys[2 * t + 1] = key;
      st_xyzz(&edge_pts[2 * t + 1], acc);
      ek1 = key;
    }
  }
  if (ek0 == kInvalidKey) edge_keys[2 * t] = kInvalidKey;
  if (ek1 == kInvalidKey) edge_keys[2 * t + 1] = kInvalidKey;
  if (dbg) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
      unsigned long long tt;
      unsigned sm;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
      asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
      dbg[3 * (t >> 5) + 1] = tt;
      dbg[3 * (t >> 5) + 2] = sm;
    }
  }
}

// ---- 5. accumulate, edge levels --------------------------------------------------------------------------
// in: L entries (key, point), keys non-decreasing over the valid entries.  Thread t folds entries
// [t*S, (t+1)*S).  With `final` set (single thread sees all entries) every group is complete.
__global__ void __launch_bounds__(128)
msm_accum_edges_kernel(const uint32_t* __restrict__ in_keys, const XYZZ* __restrict__ in_pts, size_t L,
                       XYZZ* __restrict__ buckets, uint32_t* __restrict__ out_keys, XYZZ* __restrict__ out_pts,
                       size_t num_slices, int final_level) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= num_slices) return;
  const size_t p0 = t * kSliceE;
  const size_t p1 = min(L, p0 + (final_level ? L : (size_t)kSliceE));
  uint32_t ek0 = kInvalidKey, ek1 = kInvalidKey;
  size_t pos = p0;
  bool first = true;
  while (pos < p1) {
    uint32_t key = in_keys[pos];
    if (key == kInvalidKey) { pos++; continue; }
    XYZZ acc = ld_xyzz(&in_pts[pos]);
    pos++;
    while (pos < p1) {
      const uint32_t k2 = in_keys[pos];
      if (k2 == kInvalidKey) { pos++; continue; }
      if (k2 != key) break;
      XYZZ b = ld_xyzz(&in_pts[pos]);
      xyzz_add(acc, b);
      pos++;
    }
    const bool last = (pos >= p1);
    if (final_level) {
      st_xyzz(&buckets[key], acc);
    } else if (first) {
      out_keys[2 * t] = key;
      st_xyzz(&out_pts[2 * t], acc);
      ek0 = key;
    } else if (last) {
      out_keys[2 * t + 1] = key;
      st_xyzz(&out_pts[2 * t + 1], acc);
      ek1 = key;
    } else {
      st_xyzz(&buckets[key], acc);
    }
    first = false;
  }
  if (!final_level) {
    if (ek0 == kInvalidKey) out_keys[2 * t] = kInvalidKey;
    if (ek1 == kInvalidKey) out_keys[2 * t + 1] = kInvalidKey;
  }
}

// ---- 6. bucket reduction ----------------------------------------------------------------------------------
// window sum = sum_{b=0}^{B-1} (b+1) * bucket[b]  (reference running-sum loop, src/msm/mod.zig:423-432).
// Thread (row, ch) covers buckets [ch*Lc, (ch+1)*Lc): T = sum (b-b_lo+1)*bucket[b], S = sum bucket[b],
// contribution T + b_lo*S.
__global__ void __launch_bounds__(128)
msm_chunk_sums_kernel(const XYZZ* __restrict__ buckets, uint32_t B, uint32_t Lc, uint32_t chunks_per_row,
                      size_t rows, XYZZ* __restrict__ chunk_out) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= rows * chunks_per_row) return;
  const size_t row = t / chunks_per_row;
  const uint32_t ch = (uint32_t)(t - row * chunks_per_row);
  const uint32_t b_lo = ch * Lc;
  const uint32_t b_hi = min(B, b_lo + Lc);
  const XYZZ* bk = buckets + row * B;
  XYZZ running = XYZZ::identity(), tsum = XYZZ::identity();
  for (uint32_t b = b_hi; b-- > b_lo;) {
    XYZZ v = ld_xyzz(&bk[b]);
    xyzz_add(running, v);
    xyzz_add(tsum, running);
  }
  if (b_lo) {
    XYZZ scaled = xyzz_mul_small(running, b_lo);
    xyzz_add(tsum, scaled);
  }
  st_xyzz(&chunk_out[t], tsum);
}

// One block per row (batch, window): tree over the row's chunk sums, in place in global memory.
__global__ void __launch_bounds__(256) msm_row_tree_kernel(XYZZ* chunk_sums, uint32_t chunks_per_row, XYZZ* row_out) {
  XYZZ* p = chunk_sums + (size_t)blockIdx.x * chunks_per_row;
  uint32_t active = chunks_per_row;
  while (active > 1) {
    const uint32_t half = (active + 1) >> 1;
    for (uint32_t i = threadIdx.x; i + half < active; i += blockDim.x) {
      XYZZ a = ld_xyzz_cg(&p[i]);
      XYZZ b = ld_xyzz_cg(&p[i + half]);
      xyzz_add(a, b);
      st_xyzz(&p[i], a);
    }
    __syncthreads();
    active = half;
  }
  if (threadIdx.x == 0) st_xyzz(&row_out[blockIdx.x], ld_xyzz_cg(&p[0]));
}

// result[batch] = sum_w 2^(c*w) * row[batch*W + w]  (Horner from the top window down, the reference's
// "double c times then add", src/msm/mod.zig:393-398,434).
__global__ void msm_horner_kernel(const XYZZ* __restrict__ rows, int windows, int c, size_t k, XYZZ* __restrict__ out) {
  const size_t batch = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (batch >= k) return;
  XYZZ acc = XYZZ::identity();
  for (int w = windows - 1; w >= 0; w--) {
    for (int j = 0; j < c; j++) acc = xyzz_dbl(acc);
    XYZZ r = ld_xyzz(&rows[batch * windows + w]);
    xyzz_add(acc, r);
  }
  st_xyzz(&out[batch], acc);
}

// ---- 7. normalisation ---------------------------------------------------------------------------------------
__global__ void msm_to_affine_kernel(const XYZZ* __restrict__ in, size_t k, AffineG1* __restrict__ out_xy,
                                     uint8_t* __restrict__ out_inf) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= k) return;
  AffineG1 a;
  const bool finite = xyzz_to_affine(ld_xyzz(&in[i]), a);
  st_fq(&out_xy[i].x, a.x);
  st_fq(&out_xy[i].y, a.y);
  out_inf[i] = finite ? 0 : 1;
}

__global__ void g1_sum_partials_kernel(const XYZZ* __restrict__ parts, size_t count, XYZZ* __restrict__ out) {
  XYZZ acc = XYZZ::identity();
  for (size_t i = 0; i < count; i++) {
    XYZZ p = ld_xyzz(&parts[i]);
    xyzz_add(acc, p);
  }
  st_xyzz(out, acc);
}

// Point-range shards across GPUs: publish this rank's partial (128 bytes) into every rank's NVLink mailbox,
// wait for all of them, add the rows in rank order.  One thread; the additions are 8 full XYZZ adds at most.
__global__ void g1_peer_sum_kernel(const XYZZ* __restrict__ part, const __grid_constant__ PeerDev peer, XYZZ* __restrict__ out) {
  if (threadIdx.x != 0) return;
  Fr vals[4];
  const Fr* src = reinterpret_cast<const Fr*>(part);  // XYZZ = four 32-byte elements
#pragma unroll 1
  for (int i = 0; i < 4; i++) vals[i] = src[i];
  peer_publish(peer, vals, 4);
  if (!peer_wait(peer)) return;
  XYZZ acc = XYZZ::identity();
#pragma unroll 1
  for (int g = 0; g < peer.world; g++) {
    XYZZ p = ld_xyzz_cg(reinterpret_cast<const XYZZ*>(peer_row(peer.box[peer.rank], peer.seq, g)));
    xyzz_add(acc, p);
  }
  st_xyzz(out, acc);
}

// ---- fixed-base multiplication by G = (1, 2) ---------------------------------------------------------------------
// table[w*255 + (d-1)] = d * 2^(8w) * G, affine, for w < 32, d in 1..255.
__global__ void __launch_bounds__(128) g1_fixed_table_kernel(AffineG1* __restrict__ table) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 32 * 255) return;
  const int w = t / 255, d = t % 255 + 1;
  AffineG1 gen;
  gen.x = Fq::one();
  gen.y = fp_add<FqParams>(Fq::one(), Fq::one());
  XYZZ p = xyzz_from_affine(gen);
  for (int j = 0; j < 8 * w; j++) p = xyzz_dbl(p);
  XYZZ r = xyzz_mul_small(p, (uint32_t)d);
  AffineG1 a;
  xyzz_to_affine(r, a);
  st_fq(&table[t].x, a.x);
  st_fq(&table[t].y, a.y);
}

__global__ void __launch_bounds__(128)
g1_fixed_base_kernel(const Fr* __restrict__ scalars, size_t n, const AffineG1* __restrict__ table,
                     AffineG1* __restrict__ out_xy, uint8_t* __restrict__ out_inf) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    Fr s = fp_from_mont<FrParams>(scalars[i]);
    XYZZ acc = XYZZ::identity();
#pragma unroll 1
    for (int w = 0; w < 32; w++) {
      const uint32_t d = (s.v[w >> 2] >> ((w & 3) * 8)) & 0xffu;
      if (d) xyzz_madd(acc, ld_affine(&table[w * 255 + (d - 1)]));
    }
    AffineG1 a;
    const bool finite = xyzz_to_affine(acc, a);
    st_fq(&out_xy[i].x, a.x);
    st_fq(&out_xy[i].y, a.y);
    out_inf[i] = finite ? 0 : 1;
  }
}

// scalars[i] = tau^i (Montgomery), tau = 0x12345678 (src/poly/commitment/mod.zig:189-199)
__global__ void __launch_bounds__(256) tau_powers_kernel(Fr* __restrict__ out, size_t n, uint32_t tau) {
  const Fr t = fp_from_u32<FrParams>(tau);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    Fr acc = Fr::one();
    Fr base = t;
    size_t e = i;
    while (e) {
      if (e & 1) acc = fp_mul<FrParams>(acc, base);
      base = fp_mul<FrParams>(base, base);
      e >>= 1;
    }
    out[i] = acc;
  }
}

// ---- window-level table ------------------------------------------------------------------------------------------
// pre[j*n + i] = 2^(c*j) * base[i] for j = 1..W-1 (level 0 is a copy of the bases).  One thread per base
// point of the current chunk; the W-1 new points are normalised with ONE inversion per thread
// (Montgomery's trick over their ZZZ coordinates).  tmp / prefix: (W-1) * chunk scratch, [level][point].
__global__ void __launch_bounds__(128)
msm_precompute_kernel(const AffineG1* __restrict__ bases, const uint8_t* __restrict__ inf, size_t n, size_t first,
                      size_t chunk, int c, int W, AffineG1* __restrict__ pre, XYZZ* __restrict__ tmp,
                      Fq* __restrict__ prefix) {
  const size_t ii = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (ii >= chunk) return;
  const size_t i = first + ii;
  const AffineG1 b = ld_affine(bases + i);
  st_fq(&pre[i].x, b.x);
  st_fq(&pre[i].y, b.y);
  if (inf != nullptr && inf[i] != 0) {  // identity stays the identity; its digits are never scattered
    for (int j = 1; j < W; j++) {
      st_fq(&pre[(size_t)j * n + i].x, Fq::zero());
      st_fq(&pre[(size_t)j * n + i].y, Fq::zero());
    }
    return;
  }
  XYZZ p = xyzz_from_affine(b);
  Fq prod = Fq::one();
  for (int j = 1; j < W; j++) {
    for (int t = 0; t < c; t++) p = xyzz_dbl(p);
    st_xyzz(&tmp[(size_t)(j - 1) * chunk + ii], p);
    st_fq(&prefix[(size_t)(j - 1) * chunk + ii], prod);
    prod = fq_mul(prod, p.zzz);
  }
  Fq inv = fp_inv<FqParams>(prod);  // a finite point of a prime-order group never doubles to the identity
  for (int j = W - 1; j >= 1; j--) {
    const XYZZ q = ld_xyzz_cg(&tmp[(size_t)(j - 1) * chunk + ii]);
    const uint4* pq = reinterpret_cast<const uint4*>(&prefix[(size_t)(j - 1) * chunk + ii]);
    const uint4 lo4 = __ldcg(pq), hi4 = __ldcg(pq + 1);
    Fq pf;
    pf.v[0] = lo4.x; pf.v[1] = lo4.y; pf.v[2] = lo4.z; pf.v[3] = lo4.w;
    pf.v[4] = hi4.x; pf.v[5] = hi4.y; pf.v[6] = hi4.z; pf.v[7] = hi4.w;
    const Fq izzz = fq_mul(inv, pf);  // 1 / q.zzz
    inv = fq_mul(inv, q.zzz);
    const Fq izz = fq_mul(fq_sqr(izzz), fq_sqr(q.zz));  // ZZZ^-2 * ZZ^2 = ZZ^-1
    st_fq(&pre[(size_t)j * n + i].x, fq_mul(q.x, izz));
    st_fq(&pre[(size_t)j * n + i].y, fq_mul(q.y, izzz));
  }
}

// ------------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------------
struct Arena {  // carves the context's workspace with 256-byte alignment
  char* base = nullptr;
  size_t size = 0, used = 0;
  template <class T>
  T* take(size_t count) {
    used = (used + 255) & ~size_t(255);
    T* p = reinterpret_cast<T*>(base + used);
    used += count * sizeof(T);
    return p;
  }
};

size_t edge_levels_total(size_t L0) {  // total entries over all edge levels after level 0
  size_t total = 0, L = L0;
  while (L > (size_t)kSliceE) {
    size_t slices = (L + kSliceE - 1) / kSliceE;
    L = 2 * slices;
    total += L;
  }
  return total;
}

inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

// Core: k MSMs of n points over bases[offset..offset+n), scalars on the device (k*n Fr, Montgomery).
// Leaves k XYZZ results at d_out.
int msm_core(zolt_ctx* ctx, const zolt_bases* bases, size_t offset, const Fr* d_scalars, size_t k, size_t n,
             XYZZ* d_out) {
  cudaStream_t st = ctx->stream;
  if (n == 0 || k == 0) {
    if (k) ZOLT_CUDA_TRY(cudaMemsetAsync(d_out, 0, k * sizeof(XYZZ), st));  // identity (ZZ == 0)
    return ZOLT_OK;
  }
  // Use the window-level table when the bases carry one and the MSM is large enough to fill its buckets.
  const bool pre = bases->d_pre != nullptr && n >= (size_t(1) << (bases->pre_c - 6));
  MsmGeom g = choose_geom(n);
  if (pre) {
    g.c = bases->pre_c;
    g.windows = bases->pre_w;
    g.B = 1u << (g.c - 1);
  }
  const size_t rows = k * (size_t)g.windows;       // digit rows: (batch, window)
  const size_t krows = pre ? k : rows;             // bucket sets
  const size_t num_keys = krows * g.B;
  const size_t max_entries = rows * n;
  if (num_keys >= 0xfffffff0ull || max_entries >= 0xfffffff0ull || n >= 0x7fffffffull) return ZOLT_ERR_ARG;
  if (pre && (size_t)g.windows * bases->n >= 0x7fffffffull) return ZOLT_ERR_ARG;

  // Level-0 slice length: about 128 (256 for very long lists), adjusted so that the blocks of the accumulation kernel fill
  // a whole number of waves (4 blocks of 128 threads per SM): a trailing partial wave costs a full wave of time -- at 2^21
  // points per GPU (the 8-GPU shard of 2^24) 2.81 waves of 128-entry slices became 3 waves of 120-entry ones.
  uint32_t slice_len = max_entries >= (size_t(1) << 26) ? kSlice0Big : kSlice0;
  {
    const size_t wave = (size_t)ctx->num_sms * 4 * 128;  // threads resident at once
    const size_t waves = (max_entries + (size_t)slice_len * wave - 1) / ((size_t)slice_len * wave);
    if (waves >= 1 && waves <= 64) {
      const size_t even = (max_entries + waves * wave - 1) / (waves * wave);
      if (even >= 32 && even <= slice_len) slice_len = (uint32_t)even;
    }
  }
  const size_t slices0 = (max_entries + slice_len - 1) / slice_len;
  const size_t L0 = 2 * slices0;
  const size_t edge_total = L0 + edge_levels_total(L0);
  const uint32_t Lc = kChunk;
  const uint32_t chunks_per_row = (g.B + Lc - 1) / Lc;
  const size_t num_tiles = (num_keys + kScanTile - 1) / kScanTile;

  // Two-pass bucket sort (3b) when all windows share one bucket set and the bins are worth their set-up
  const bool two_pass = pre && k == 1 && g.B >= 4096 && g.B <= (1u << 20) && g.windows <= kBinMaxWindows && n >= (size_t(1) << 18) && ctx->msm_sort != 1;
  const uint32_t nbins = g.B >> kBinSubBits;
  const uint32_t bin_grid = (uint32_t)ctx->num_sms;           // blocks of the two partition kernels: one per SM at 1024 threads (two of 512 measured slower)
  const size_t matrix_n = two_pass ? (size_t)nbins * bin_grid : 0;
  const size_t matrix_tiles = (matrix_n + kScanTile - 1) / kScanTile;
  const size_t max_chunks = two_pass ? max_entries / kChunkRecs + nbins + 1 : 0;

  size_t bytes = 0;
  bytes += align256(max_entries * 4) * (two_pass ? 3 : 2);    // digits (or 8-byte records), sorted
  bytes += align256((matrix_n + 1) * 4) * 2 + align256((matrix_tiles + 1) * 4) + align256(((size_t)nbins + 1) * 4);
  bytes += align256(num_keys * 4) + align256((num_keys + 1) * 4);  // counts/cursor, offsets
  bytes += align256((num_tiles + 1) * 4);
  bytes += align256(num_keys * sizeof(XYZZ));                 // buckets
  bytes += align256(edge_total * 4) + align256(edge_total * sizeof(XYZZ));
  const uint32_t tree_parts = chunks_per_row > 1024 ? chunks_per_row / 256 : 1;  // two-stage row tree
  bytes += align256(krows * chunks_per_row * sizeof(XYZZ)) + align256(krows * sizeof(XYZZ));
  bytes += align256(krows * tree_parts * sizeof(XYZZ));
  bytes += 4096;
  if (ctx->ws_bytes < bytes) {
    if (ctx->ws) {
      ZOLT_CUDA_TRY(cudaStreamSynchronize(st));
      ZOLT_CUDA_TRY(cudaFree(ctx->ws));
      ctx->ws = nullptr;
      ctx->ws_bytes = 0;
    }
    ZOLT_CUDA_TRY(cudaMalloc(&ctx->ws, bytes));
    ctx->ws_bytes = bytes;
  }
  Arena ar;
  ar.base = static_cast<char*>(ctx->ws);
  ar.size = bytes;
  uint32_t* digits = ar.take<uint32_t>(max_entries * (two_pass ? 2 : 1));
  uint32_t* sorted = ar.take<uint32_t>(max_entries);
  uint32_t* matrix = ar.take<uint32_t>(matrix_n + 1);
  uint32_t* matrix_off = ar.take<uint32_t>(matrix_n + 1);
  uint32_t* matrix_tiles_buf = ar.take<uint32_t>(matrix_tiles + 1);
  uint32_t* chunk_base = ar.take<uint32_t>((size_t)nbins + 1);
  uint32_t* counts = ar.take<uint32_t>(num_keys);
  uint32_t* offsets = ar.take<uint32_t>(num_keys + 1);
  uint32_t* tiles = ar.take<uint32_t>(num_tiles + 1);
  XYZZ* buckets = ar.take<XYZZ>(num_keys);
  uint32_t* edge_keys = ar.take<uint32_t>(edge_total);
  XYZZ* edge_pts = ar.take<XYZZ>(edge_total);
  XYZZ* chunk_sums = ar.take<XYZZ>(krows * chunks_per_row);
  XYZZ* row_sums = ar.take<XYZZ>(krows);
  XYZZ* part_sums = ar.take<XYZZ>(krows * tree_parts);

  int status = ZOLT_OK;
  auto fail = [&](cudaError_t e, const char* what) {
    fprintf(stderr, "[zolt_gpu] msm: %s -> %s\n", what, cudaGetErrorString(e));
    status = ZOLT_ERR_CUDA;
  };
#define MSM_CHECK(expr)                              \
  do {                                               \
    cudaError_t _e = (expr);                         \
    if (_e != cudaSuccess) { fail(_e, #expr); goto done; } \
  } while (0)

  {
    const AffineG1* d_pts = pre ? bases->d_pre : bases->d_xy + offset;
    const uint8_t* d_inf = bases->d_inf ? bases->d_inf + offset : nullptr;
    const unsigned sm = (unsigned)ctx->num_sms;

    MSM_CHECK(cudaMemsetAsync(counts, 0, num_keys * 4, st));
    MSM_CHECK(cudaMemsetAsync(buckets, 0, num_keys * sizeof(XYZZ), st));  // identity everywhere
    if (two_pass) {
      uint2* recs = reinterpret_cast<uint2*>(digits);
      const size_t bin_smem = (size_t)nbins * 4;
      // partition kernel: cursor | cnt | loff | one tile of records (8 B) and their bins (2 B)
      const size_t part_smem = 3 * bin_smem + (size_t)kBinThreads * g.windows * (8 + 2) + 16;
      static std::atomic<unsigned long long> attr_devs{0};  // > 48 KB of dynamic shared memory needs an opt-in per device
      if (!(attr_devs.load() & (1ull << (ctx->device & 63)))) {
        MSM_CHECK(cudaFuncSetAttribute(msm_bin_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        MSM_CHECK(cudaFuncSetAttribute(msm_bin_partition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        MSM_CHECK(cudaFuncSetAttribute(msm_chunk_place_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kChunkRecs * 6));
        attr_devs.fetch_or(1ull << (ctx->device & 63));
      }
      int ps = prof_begin(ctx, PROF_MSM_DIGITS);
      msm_bin_hist_kernel<<<bin_grid, kBinThreads, bin_smem, st>>>(d_scalars, d_inf, n, g, nbins, matrix);
      scan_tile_sums_kernel<<<(unsigned)matrix_tiles, kScanThreads, 0, st>>>(matrix, matrix_n, matrix_tiles_buf);
      scan_tile_offsets_kernel<<<1, kScanThreads, 0, st>>>(matrix_tiles_buf, matrix_tiles);
      scan_apply_kernel<<<(unsigned)matrix_tiles, kScanThreads, 0, st>>>(matrix, matrix_n, matrix_tiles_buf, matrix_off);
      msm_bin_partition_kernel<<<bin_grid, kBinThreads, part_smem, st>>>(d_scalars, d_inf, n, g, nbins, matrix_off, bases->n, offset, recs);
      msm_chunk_index_kernel<<<1, kScanThreads, 0, st>>>(matrix_off, nbins, bin_grid, chunk_base);
      prof_end(ctx, ps);
      ps = prof_begin(ctx, PROF_MSM_SCAN);
      msm_chunk_count_kernel<<<(unsigned)max_chunks, kChunkThreads, 0, st>>>(recs, chunk_base, matrix_off, nbins, bin_grid, counts);
      scan_tile_sums_kernel<<<(unsigned)num_tiles, kScanThreads, 0, st>>>(counts, num_keys, tiles);
      scan_tile_offsets_kernel<<<1, kScanThreads, 0, st>>>(tiles, num_tiles);
      scan_apply_kernel<<<(unsigned)num_tiles, kScanThreads, 0, st>>>(counts, num_keys, tiles, offsets);
      prof_end(ctx, ps);
      MSM_CHECK(cudaMemcpyAsync(counts, offsets, num_keys * 4, cudaMemcpyDeviceToDevice, st));
      ps = prof_begin(ctx, PROF_MSM_SCATTER);
      msm_chunk_place_kernel<<<(unsigned)max_chunks, kChunkThreads, kChunkRecs * 6, st>>>(recs, chunk_base, matrix_off, nbins, bin_grid, counts, sorted);
      prof_end(ctx, ps);
      ctx->launches += 11;
      MSM_CHECK(cudaGetLastError());
    } else {
    {
      size_t blocks = (k * n + 255) / 256;
      if (blocks > (size_t)sm * 16) blocks = (size_t)sm * 16;
      const int ps = prof_begin(ctx, PROF_MSM_DIGITS);
      msm_digits_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_scalars, d_inf, n, k, g, pre ? 1 : 0, digits, counts);
      prof_end(ctx, ps);
      ctx->launches++;
    }
    int ps = prof_begin(ctx, PROF_MSM_SCAN);
    scan_tile_sums_kernel<<<(unsigned)num_tiles, kScanThreads, 0, st>>>(counts, num_keys, tiles);
    scan_tile_offsets_kernel<<<1, kScanThreads, 0, st>>>(tiles, num_tiles);
    scan_apply_kernel<<<(unsigned)num_tiles, kScanThreads, 0, st>>>(counts, num_keys, tiles, offsets);
    prof_end(ctx, ps);
    ctx->launches += 3;
    // counts is reused as the scatter cursor, pre-loaded with the offsets: one random access per entry
    MSM_CHECK(cudaMemcpyAsync(counts, offsets, num_keys * 4, cudaMemcpyDeviceToDevice, st));
    {
      ps = prof_begin(ctx, PROF_MSM_SCATTER);
      {
        // 2-D grid: y walks the digit rows, x the points of a row; about 16 blocks per SM in total
        const unsigned gy = (unsigned)(rows < 1024 ? rows : 1024);
        size_t gx = ((size_t)sm * 16 + gy - 1) / gy;
        const size_t gx_max = (n + 255) / 256;
        if (gx > gx_max) gx = gx_max;
        if (gx == 0) gx = 1;
        msm_scatter_kernel<<<dim3((unsigned)gx, gy), 256, 0, st>>>(digits, n, rows, g, pre ? 1 : 0, bases->n, offset, counts, sorted);
      }
      prof_end(ctx, ps);
      ctx->launches++;
    }
    }
    // level 0
    int ps = prof_begin(ctx, PROF_MSM_ACCUM0);
    msm_accum_points_kernel<<<(unsigned)((slices0 + 127) / 128), 128, 0, st>>>(
        sorted, offsets, (uint32_t)num_keys, d_pts, buckets, edge_keys, edge_pts, slices0, slice_len, ctx->msm_dbg);
    prof_end(ctx, ps);
    ctx->launches++;
    // edge levels
    ps = prof_begin(ctx, PROF_MSM_EDGES);
    {
      size_t L = L0, in_off = 0;
      for (;;) {
        const bool final_level = (L <= (size_t)kSliceE);
        const size_t slices = final_level ? 1 : (L + kSliceE - 1) / kSliceE;
        const size_t out_off = in_off + L;
        msm_accum_edges_kernel<<<(unsigned)((slices + 127) / 128), 128, 0, st>>>(
            edge_keys + in_off, edge_pts + in_off, L, buckets, edge_keys + out_off, edge_pts + out_off, slices,
            final_level ? 1 : 0);
        ctx->launches++;
        if (final_level) break;
        in_off = out_off;
        L = 2 * slices;
      }
    }
    prof_end(ctx, ps);
    // bucket reduction
    ps = prof_begin(ctx, PROF_MSM_REDUCE);
    {
      const size_t threads = krows * chunks_per_row;
      msm_chunk_sums_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(buckets, g.B, Lc, chunks_per_row, krows,
                                                                             chunk_sums);
      auto tree_threads = [](uint32_t chunks) {
        unsigned tb = ((chunks + 1) / 2 + 31) / 32 * 32;
        return tb > 256 ? 256u : (tb ? tb : 32u);
      };
      if (tree_parts > 1) {
        // a row's chunk sums only need to be added up: first 256 at a time over many blocks, then the parts
        msm_row_tree_kernel<<<(unsigned)(krows * tree_parts), tree_threads(256), 0, st>>>(chunk_sums, 256, part_sums);
        msm_row_tree_kernel<<<(unsigned)krows, tree_threads(tree_parts), 0, st>>>(part_sums, tree_parts, row_sums);
        ctx->launches++;
      } else {
        msm_row_tree_kernel<<<(unsigned)krows, tree_threads(chunks_per_row), 0, st>>>(chunk_sums, chunks_per_row, row_sums);
      }
      msm_horner_kernel<<<(unsigned)((k + 31) / 32), 32, 0, st>>>(row_sums, pre ? 1 : g.windows, pre ? 0 : g.c, k, d_out);
      ctx->launches += 3;
    }
    prof_end(ctx, ps);
    MSM_CHECK(cudaGetLastError());
  }
done:
#undef MSM_CHECK
  return status;
}

int staging_for(zolt_ctx* ctx, size_t bytes, Fr** out) {
  if (ctx->ws_in_bytes < bytes) {
    if (ctx->ws_in) {
      ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
      ZOLT_CUDA_TRY(cudaFree(ctx->ws_in));
      ctx->ws_in = nullptr;
      ctx->ws_in_bytes = 0;
    }
    ZOLT_CUDA_TRY(cudaMalloc(&ctx->ws_in, bytes ? bytes : 32));
    ctx->ws_in_bytes = bytes ? bytes : 32;
  }
  *out = static_cast<Fr*>(ctx->ws_in);
  return ZOLT_OK;
}

int finish_affine(zolt_ctx* ctx, const XYZZ* d_xyzz, size_t k, uint64_t* out_xy, uint8_t* out_inf) {
  AffineG1* d_xy = nullptr;
  uint8_t* d_inf = nullptr;
  ZOLT_CUDA_TRY(pool_alloc(ctx, k * sizeof(AffineG1), (void**)&d_xy));
  ZOLT_CUDA_TRY(pool_alloc(ctx, k, (void**)&d_inf));
  const int ps = prof_begin(ctx, PROF_MSM_AFFINE);
  msm_to_affine_kernel<<<(unsigned)((k + 31) / 32), 32, 0, ctx->stream>>>(d_xyzz, k, d_xy, d_inf);
  prof_end(ctx, ps);
  ctx->launches++;
  ZOLT_CUDA_TRY(cudaGetLastError());
  ZOLT_CUDA_TRY(cudaMemcpyAsync(out_xy, d_xy, k * sizeof(AffineG1), cudaMemcpyDeviceToHost, ctx->stream));
  ZOLT_CUDA_TRY(cudaMemcpyAsync(out_inf, d_inf, k, cudaMemcpyDeviceToHost, ctx->stream));
  ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  pool_release(ctx, d_xy, k * sizeof(AffineG1));
  pool_release(ctx, d_inf, k);
  return ZOLT_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" int zolt_msm_bases_upload(zolt_ctx* ctx, const uint64_t* xy, const uint8_t* inf, size_t n, zolt_bases** out) {
  if (!ctx || !out || (n && !xy)) return ZOLT_ERR_ARG;
  *out = nullptr;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  zolt_bases* b = new (std::nothrow) zolt_bases();
  if (!b) return ZOLT_ERR_ARG;
  b->ctx = ctx;
  b->n = n;
  OnFail undo([&] { zolt_msm_bases_free(b); });
  ZOLT_CUDA_TRY(cudaMalloc((void**)&b->d_xy, (n ? n : 1) * sizeof(AffineG1)));
  ZOLT_CUDA_TRY(cudaMemcpyAsync(b->d_xy, xy, n * sizeof(AffineG1), cudaMemcpyHostToDevice, ctx->stream));
  bool any_inf = false;
  if (inf)
    for (size_t i = 0; i < n; i++)
      if (inf[i]) { any_inf = true; break; }
  if (any_inf) {
    ZOLT_CUDA_TRY(cudaMalloc((void**)&b->d_inf, n));
    ZOLT_CUDA_TRY(cudaMemcpyAsync(b->d_inf, inf, n, cudaMemcpyHostToDevice, ctx->stream));
  }
  ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  undo.dismiss();
  *out = b;
  return ZOLT_OK;
}

extern "C" int zolt_msm_bases_free(zolt_bases* b) {
  if (!b) return ZOLT_ERR_ARG;
  cudaSetDevice(b->ctx->device);
  cudaStreamSynchronize(b->ctx->stream);
  if (b->d_xy && !b->borrowed) cudaFree(b->d_xy);
  if (b->d_inf && !b->borrowed) cudaFree(b->d_inf);
  if (b->d_pre) cudaFree(b->d_pre);
  delete b;
  return ZOLT_OK;
}

extern "C" int zolt_msm_bases_precompute(zolt_bases* b, int window_bits) {
  if (!b) return ZOLT_ERR_ARG;
  zolt_ctx* ctx = b->ctx;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  // measured on B200 (tools/gpu_phases.py): 2^24 points 51.5 -> 43.4 ms at c = 20 (c = 22 saves additions but its
  // counting-sort scatter spreads over a larger array), 2^21 9.0 -> 6.9 ms at c = 20, 2^18 3.8 -> 2.2 ms at c = 16
  if (window_bits == 0) window_bits = b->n >= (size_t(1) << 20) ? 20 : (b->n >= (size_t(1) << 15) ? 16 : 0);
  if (window_bits == 0 || b->n == 0) return ZOLT_OK;  // too small to profit: MSMs keep the per-window path
  if (window_bits < 8 || window_bits > 26) return ZOLT_ERR_ARG;
  const int W = (255 + window_bits - 1) / window_bits;
  if ((size_t)W * b->n >= 0x7fffffffull) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (b->d_pre) {
    cudaFree(b->d_pre);
    b->d_pre = nullptr;
  }
  AffineG1* pre = nullptr;
  XYZZ* tmp = nullptr;
  Fq* prefix = nullptr;
  OnFail undo([&] {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(pre);
    cudaFree(tmp);
    cudaFree(prefix);
  });
  ZOLT_CUDA_TRY(cudaMalloc((void**)&pre, (size_t)W * b->n * sizeof(AffineG1)));
  const size_t chunk_max = size_t(1) << 20;
  const size_t chunk_alloc = b->n < chunk_max ? b->n : chunk_max;
  ZOLT_CUDA_TRY(cudaMalloc((void**)&tmp, (size_t)(W - 1) * chunk_alloc * sizeof(XYZZ)));
  ZOLT_CUDA_TRY(cudaMalloc((void**)&prefix, (size_t)(W - 1) * chunk_alloc * sizeof(Fq)));
  for (size_t first = 0; first < b->n; first += chunk_alloc) {
    const size_t chunk = b->n - first < chunk_alloc ? b->n - first : chunk_alloc;
    msm_precompute_kernel<<<(unsigned)((chunk + 127) / 128), 128, 0, ctx->stream>>>(b->d_xy, b->d_inf, b->n, first, chunk,
                                                                                   window_bits, W, pre, tmp, prefix);
    ctx->launches++;
  }
  ZOLT_CUDA_TRY(cudaGetLastError());
  ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  undo.dismiss();
  cudaFree(tmp);
  cudaFree(prefix);
  b->d_pre = pre;
  b->pre_c = window_bits;
  b->pre_w = W;
  return ZOLT_OK;
}

extern "C" int zolt_msm_bases_len(const zolt_bases* b, size_t* n) {
  if (!b || !n) return ZOLT_ERR_ARG;
  *n = b->n;
  return ZOLT_OK;
}

extern "C" int zolt_msm_bases_download(const zolt_bases* b, size_t offset, size_t n, uint64_t* xy, uint8_t* inf) {
  if (!b || !xy || offset + n > b->n) return ZOLT_ERR_ARG;
  zolt_ctx* ctx = b->ctx;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  ZOLT_CUDA_TRY(cudaMemcpyAsync(xy, b->d_xy + offset, n * sizeof(AffineG1), cudaMemcpyDeviceToHost, ctx->stream));
  if (inf) {
    if (b->d_inf) ZOLT_CUDA_TRY(cudaMemcpyAsync(inf, b->d_inf + offset, n, cudaMemcpyDeviceToHost, ctx->stream));
    else memset(inf, 0, n);
  }
  ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return ZOLT_OK;
}

extern "C" int zolt_msm_partial_dev(zolt_ctx* ctx, const zolt_bases* bases, size_t offset, const void* d_scalars,
                                    size_t n, void* d_out) {
  if (!ctx || !bases || !d_out || (n && !d_scalars) || offset + n > bases->n) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  return msm_core(ctx, bases, offset, (const Fr*)d_scalars, 1, n, (XYZZ*)d_out);
}

extern "C" int zolt_g1_sum_partials(zolt_ctx* ctx, const void* d_partials, size_t count, uint64_t out_xy[8],
                                    uint8_t* out_inf) {
  if (!ctx || !d_partials || !out_xy || !out_inf) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  XYZZ* d_sum = nullptr;
  ZOLT_CUDA_TRY(pool_alloc(ctx, sizeof(XYZZ), (void**)&d_sum));
  g1_sum_partials_kernel<<<1, 1, 0, ctx->stream>>>((const XYZZ*)d_partials, count, d_sum);
  ctx->launches++;
  int s = finish_affine(ctx, d_sum, 1, out_xy, out_inf);
  pool_release(ctx, d_sum, sizeof(XYZZ));
  return s;
}

extern "C" int zolt_msm_sharded_dev(zolt_ctx* ctx, const zolt_bases* bases, size_t offset, const void* d_scalars, size_t n,
                                    zolt_peer_box* box, uint64_t out_xy[8], uint8_t* out_inf) {
  if (!ctx || !bases || !box || !box->connected || box->ctx != ctx || !out_xy || !out_inf || (n && !d_scalars) || offset + n > bases->n)
    return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  XYZZ* d_res = nullptr;
  ZOLT_CUDA_TRY(pool_alloc(ctx, 2 * sizeof(XYZZ), (void**)&d_res));
  int s = msm_core(ctx, bases, offset, (const Fr*)d_scalars, 1, n, d_res);
  if (s == ZOLT_OK) {
    g1_peer_sum_kernel<<<1, 32, 0, ctx->stream>>>(d_res, peer_next(box), d_res + 1);
    ctx->launches++;
    s = finish_affine(ctx, d_res + 1, 1, out_xy, out_inf);
    if (s == ZOLT_OK) s = peer_check(box);
  }
  pool_release(ctx, d_res, 2 * sizeof(XYZZ));
  return s;
}

extern "C" int zolt_msm_dev(zolt_ctx* ctx, const zolt_bases* bases, size_t offset, const void* d_scalars, size_t n,
                            uint64_t out_xy[8], uint8_t* out_inf) {
  if (!ctx || !bases || !out_xy || !out_inf || (n && !d_scalars) || offset + n > bases->n) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  XYZZ* d_res = nullptr;
  ZOLT_CUDA_TRY(pool_alloc(ctx, sizeof(XYZZ), (void**)&d_res));
  int s = msm_core(ctx, bases, offset, (const Fr*)d_scalars, 1, n, d_res);
  if (s == ZOLT_OK) s = finish_affine(ctx, d_res, 1, out_xy, out_inf);
  pool_release(ctx, d_res, sizeof(XYZZ));
  return s;
}

// sum_i scalars[i] * bases[offset + i] for HOST scalars, left on the device as one XYZZ point (d_out[0]; d_out needs room
// for 4 points).  Large inputs: the scalars take ~10 ms per 2^24 over PCIe, so the points are split in ranges
// (ParallelMSM's split, src/msm/mod.zig:608-653, used here for overlap): every range but the first is copied on a second
// stream while the range before it is multiplied, then the partial sums are added.  Same group element either way.
static int msm_host_partial(zolt_ctx* ctx, const zolt_bases* bases, size_t offset, const uint64_t* scalars, size_t n, XYZZ* d_out) {
  Fr* d_s = nullptr;
  ZOLT_TRY(staging_for(ctx, n * sizeof(Fr), &d_s));
  if (n < (size_t(1) << 22)) {
    ZOLT_CUDA_TRY(cudaMemcpyAsync(d_s, scalars, n * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    return msm_core(ctx, bases, offset, d_s, 1, n, d_out);
  }
  if (!ctx->copy_stream) {
    ZOLT_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    ZOLT_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_gate, cudaEventDisableTiming));
    ZOLT_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_copied, cudaEventDisableTiming));
    ZOLT_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_copied2, cudaEventDisableTiming));
  }
  // Point ranges whose uploads hide behind the MSM of the range before: the bus moves a scalar ~4x faster than the GPU
  // multiplies it, so a small first range (the only upload nobody hides) followed by growing ones -- 1/8, 3/8, 4/8 from 2^23
  // points (e2e 2^24: 42.9 -> 40.9 ms), two halves below that (every extra range pays the per-call fixed costs once more).
  size_t part[3];
  int parts;
  if (n >= (size_t(1) << 23)) {
    parts = 3;
    part[0] = n / 8;
    part[1] = 3 * (n / 8);
    part[2] = n - part[0] - part[1];
  } else {
    parts = 2;
    part[0] = (n + 1) / 2;
    part[1] = n - part[0];
  }
  ZOLT_CUDA_TRY(cudaMemcpyAsync(d_s, scalars, part[0] * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
  ZOLT_CUDA_TRY(cudaEventRecord(ctx->ev_gate, ctx->stream));  // the later copies queue behind the first on the bus
  ZOLT_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_gate, 0));
  // Each range's MSM is launched BEFORE the next range's copy is issued: from pageable host memory cudaMemcpyAsync blocks
  // the calling thread, and that time should pass while the GPU multiplies.
  size_t done = 0;
  for (int i = 0; i < parts; i++) {
    if (i > 0) ZOLT_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, i == 1 ? ctx->ev_copied : ctx->ev_copied2, 0));
    ZOLT_TRY(msm_core(ctx, bases, offset + done, d_s + done, 1, part[i], d_out + 1 + i));
    done += part[i];
    if (i + 1 < parts) {
      ZOLT_CUDA_TRY(cudaMemcpyAsync(d_s + done, scalars + 4 * done, part[i + 1] * sizeof(Fr), cudaMemcpyHostToDevice, ctx->copy_stream));
      ZOLT_CUDA_TRY(cudaEventRecord(i == 0 ? ctx->ev_copied : ctx->ev_copied2, ctx->copy_stream));
    }
  }
  g1_sum_partials_kernel<<<1, 1, 0, ctx->stream>>>(d_out + 1, parts, d_out);
  ctx->launches++;
  ZOLT_CUDA_TRY(cudaGetLastError());
  return ZOLT_OK;
}

extern "C" int zolt_msm(zolt_ctx* ctx, const zolt_bases* bases, size_t offset, const uint64_t* scalars, size_t n,
                        uint64_t out_xy[8], uint8_t* out_inf) {
  if (!ctx || !bases || !out_xy || !out_inf || (n && !scalars) || offset + n > bases->n) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  XYZZ* d_res = nullptr;
  ZOLT_CUDA_TRY(pool_alloc(ctx, 4 * sizeof(XYZZ), (void**)&d_res));
  int s = msm_host_partial(ctx, bases, offset, scalars, n, d_res);
  if (s == ZOLT_OK) s = finish_affine(ctx, d_res, 1, out_xy, out_inf);
  pool_release(ctx, d_res, 4 * sizeof(XYZZ));
  return s;
}

// The un-normalised partial sum from HOST scalars, left on the device (128 bytes at d_out): zolt_msm_partial_dev with the
// upload overlapped as in zolt_msm.  Used by the one-process multi-device MSM (multi.cu).
extern "C" int zolt_msm_partial(zolt_ctx* ctx, const zolt_bases* bases, size_t offset, const uint64_t* scalars, size_t n, void* d_out) {
  if (!ctx || !bases || !d_out || (n && !scalars) || offset + n > bases->n) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  XYZZ* d_res = nullptr;
  ZOLT_CUDA_TRY(pool_alloc(ctx, 4 * sizeof(XYZZ), (void**)&d_res));
  int s = msm_host_partial(ctx, bases, offset, scalars, n, d_res);
  if (s == ZOLT_OK && cudaMemcpyAsync(d_out, d_res, sizeof(XYZZ), cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) s = ZOLT_ERR_CUDA;
  pool_release(ctx, d_res, 4 * sizeof(XYZZ));
  return s;
}

// zolt_msm_sharded_dev for HOST scalars (this rank's range): upload overlapped as in zolt_msm, then the mailbox exchange.
extern "C" int zolt_msm_sharded(zolt_ctx* ctx, const zolt_bases* bases, size_t offset, const uint64_t* scalars, size_t n,
                                zolt_peer_box* box, uint64_t out_xy[8], uint8_t* out_inf) {
  if (!ctx || !bases || !box || !box->connected || box->ctx != ctx || !out_xy || !out_inf || (n && !scalars) || offset + n > bases->n)
    return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  XYZZ* d_res = nullptr;
  ZOLT_CUDA_TRY(pool_alloc(ctx, 4 * sizeof(XYZZ), (void**)&d_res));
  int s = msm_host_partial(ctx, bases, offset, scalars, n, d_res);
  if (s == ZOLT_OK) {
    g1_peer_sum_kernel<<<1, 32, 0, ctx->stream>>>(d_res, peer_next(box), d_res + 3);
    ctx->launches++;
    s = finish_affine(ctx, d_res + 3, 1, out_xy, out_inf);
    if (s == ZOLT_OK) s = peer_check(box);
  }
  pool_release(ctx, d_res, 4 * sizeof(XYZZ));
  return s;
}

extern "C" int zolt_msm_batch(zolt_ctx* ctx, const zolt_bases* bases, const uint64_t* scalars, size_t k, size_t n,
                              uint64_t* out_xy, uint8_t* out_inf) {
  if (!ctx || !bases || !out_xy || !out_inf || (n && k && !scalars) || n > bases->n) return ZOLT_ERR_ARG;
  if (k == 0) return ZOLT_OK;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  Fr* d_s = nullptr;
  XYZZ* d_res = nullptr;
  ZOLT_TRY(staging_for(ctx, n * k * sizeof(Fr), &d_s));
  ZOLT_CUDA_TRY(pool_alloc(ctx, k * sizeof(XYZZ), (void**)&d_res));  // from the context's buffer cache, like the workspace
  OnFail undo([&] { pool_release(ctx, d_res, k * sizeof(XYZZ)); });
  ZOLT_CUDA_TRY(cudaMemcpyAsync(d_s, scalars, n * k * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
  int s = msm_core(ctx, bases, 0, d_s, k, n, d_res);
  if (s == ZOLT_OK) s = finish_affine(ctx, d_res, k, out_xy, out_inf);  // synchronises the stream
  return s;  // `undo` hands the result buffer back on every path
}

static int fixed_base_dev(zolt_ctx* ctx, const Fr* d_scalars, size_t n, zolt_bases** out) {
  zolt_bases* b = new (std::nothrow) zolt_bases();
  if (!b) return ZOLT_ERR_ARG;
  b->ctx = ctx;
  b->n = n;
  AffineG1* d_table = nullptr;
  OnFail undo([&] {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_table);
    zolt_msm_bases_free(b);
  });
  ZOLT_CUDA_TRY(cudaMalloc((void**)&b->d_xy, (n ? n : 1) * sizeof(AffineG1)));
  ZOLT_CUDA_TRY(cudaMalloc((void**)&b->d_inf, n ? n : 1));
  ZOLT_CUDA_TRY(cudaMalloc((void**)&d_table, 32 * 255 * sizeof(AffineG1)));
  g1_fixed_table_kernel<<<(32 * 255 + 127) / 128, 128, 0, ctx->stream>>>(d_table);
  ctx->launches++;
  if (n) {
    size_t blocks = (n + 127) / 128;
    if (blocks > (size_t)ctx->num_sms * 16) blocks = (size_t)ctx->num_sms * 16;
    g1_fixed_base_kernel<<<(unsigned)blocks, 128, 0, ctx->stream>>>(d_scalars, n, d_table, b->d_xy, b->d_inf);
    ctx->launches++;
  }
  ZOLT_CUDA_TRY(cudaGetLastError());
  ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  undo.dismiss();
  cudaFree(d_table);
  *out = b;
  return ZOLT_OK;
}

extern "C" int zolt_g1_fixed_base_mul(zolt_ctx* ctx, const uint64_t* scalars, size_t n, zolt_bases** out) {
  if (!ctx || !out || (n && !scalars)) return ZOLT_ERR_ARG;
  *out = nullptr;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  Fr* d_s = nullptr;
  ZOLT_CUDA_TRY(cudaMalloc((void**)&d_s, (n ? n : 1) * sizeof(Fr)));
  ZOLT_CUDA_TRY(cudaMemcpyAsync(d_s, scalars, n * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
  int s = fixed_base_dev(ctx, d_s, n, out);
  cudaFree(d_s);
  return s;
}

// ---- HyperKZG.open (src/poly/commitment/mod.zig:261-324) -----------------------------------------------------
// One level of the fold: q[j] = hi[j] - lo[j] (the quotient that gets committed) and, in place,
// cur[j] = (1 - r) lo[j] + r hi[j] = lo[j] + r q[j] (the same canonical element; multiplier table as in the sumcheck bind).
__global__ void __launch_bounds__(256)
hk_quotient_fold_kernel(Fr* __restrict__ cur, size_t half, Fr* __restrict__ q, const __grid_constant__ FixedMulTable rt) {
  for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < half; j += (size_t)gridDim.x * blockDim.x) {
    const Fr lo = cur[j], hi = cur[j + half];
    q[j] = fp_sub<FrParams>(hi, lo);
    cur[j] = fp_bind_fixed<FrParams>(lo, hi, rt);
  }
}

// Levels whose quotient has at most kOpenBatchN entries are committed together: their MSMs are all launch latency
// (~1.5 ms each for ~45 kernels), so the quotients are parked as zero-padded rows of one [levels][nb] matrix and
// committed by ONE batched MSM over the first nb powers (zero digits are skipped; BatchMSM shape, src/msm/mod.zig:551-563).
constexpr size_t kOpenBatchN = 4096;

static int hyperkzg_open_core(zolt_ctx* ctx, const zolt_bases* params, Fr* d_cur /* owned copy, n */, size_t n, const uint64_t* point,
                              int num_vars, uint64_t* quotients_xy, uint8_t* quotients_inf, int* n_quotients, uint64_t final_eval[4]) {
  cudaStream_t st = ctx->stream;
  Fr *d_q = nullptr, *d_qb = nullptr;
  XYZZ* d_res = nullptr;
  const size_t q_bytes = (n / 2 ? n / 2 : 1) * sizeof(Fr);
  const size_t qb_bytes = 16 * kOpenBatchN * sizeof(Fr);  // log2(kOpenBatchN) + 1 = 13 rows at most
  ZOLT_CUDA_TRY(pool_alloc(ctx, q_bytes, (void**)&d_q));
  ZOLT_CUDA_TRY(pool_alloc(ctx, qb_bytes, (void**)&d_qb));
  ZOLT_CUDA_TRY(pool_alloc(ctx, (size_t)(num_vars ? num_vars : 1) * sizeof(XYZZ), (void**)&d_res));
  int done = 0, status = ZOLT_OK;
  int batch_first = -1, batch_rows = 0;
  size_t batch_n = 0;
  size_t len = n;
  for (int i = 0; i < num_vars && status == ZOLT_OK; i++) {
    const size_t half = len / 2;
    if (half == 0) break;  // fewer entries than 2^num_vars: the reference stops folding here (mod.zig:289)
    Fr r;
    for (int j = 0; j < 4; j++) {
      r.v[2 * j] = (uint32_t)point[4 * i + j];
      r.v[2 * j + 1] = (uint32_t)(point[4 * i + j] >> 32);
    }
    const size_t m = half < params->n ? half : params->n;  // commit(): n = min(len(evals), len(powers)) (mod.zig:241-249)
    Fr* q_dst = d_q;
    const bool batched = half <= kOpenBatchN && params->n >= half;  // an SRS shorter than the quotient: per-level path
    if (batched) {
      if (batch_first < 0) {
        batch_first = done;
        batch_n = m;
        if (cudaMemsetAsync(d_qb, 0, qb_bytes, st) != cudaSuccess) { status = ZOLT_ERR_CUDA; break; }
      }
      q_dst = d_qb + (size_t)batch_rows * batch_n;  // rows shorter than batch_n stay zero-padded
      batch_rows++;
    }
    size_t blocks = (half + 255) / 256;
    if (blocks > (size_t)ctx->num_sms * 8) blocks = (size_t)ctx->num_sms * 8;
    hk_quotient_fold_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_cur, half, q_dst, fixed_mul_table<FrParams>(r));
    ctx->launches++;
    if (!batched) status = msm_core(ctx, params, 0, d_q, 1, m, d_res + done);
    done++;
    len = half;
  }
  if (status == ZOLT_OK && batch_rows) status = msm_core(ctx, params, 0, d_qb, (size_t)batch_rows, batch_n, d_res + batch_first);
  if (status == ZOLT_OK && done) status = finish_affine(ctx, d_res, (size_t)done, quotients_xy, quotients_inf);
  if (status == ZOLT_OK) {
    if (len > 0) {
      cudaError_t e = cudaMemcpyAsync(final_eval, d_cur, sizeof(Fr), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) status = ZOLT_ERR_CUDA;
    } else {
      memset(final_eval, 0, 32);
    }
  }
  if (n_quotients) *n_quotients = done;
  pool_release(ctx, d_q, q_bytes);
  pool_release(ctx, d_qb, qb_bytes);
  pool_release(ctx, d_res, (size_t)(num_vars ? num_vars : 1) * sizeof(XYZZ));
  return status;
}

extern "C" int zolt_hyperkzg_open_dev(zolt_ctx* ctx, const zolt_bases* params, const void* d_evals, size_t n, const uint64_t* point,
                                      int num_vars, uint64_t* quotients_xy, uint8_t* quotients_inf, int* n_quotients,
                                      uint64_t final_eval[4]) {
  if (!ctx || !params || !final_eval || num_vars < 0 || (n && !d_evals) || (num_vars && (!point || !quotients_xy || !quotients_inf)))
    return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  Fr* d_cur = nullptr;
  ZOLT_CUDA_TRY(pool_alloc(ctx, (n ? n : 1) * sizeof(Fr), (void**)&d_cur));
  if (n) ZOLT_CUDA_TRY(cudaMemcpyAsync(d_cur, d_evals, n * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
  const int s = hyperkzg_open_core(ctx, params, d_cur, n, point, num_vars, quotients_xy, quotients_inf, n_quotients, final_eval);
  pool_release(ctx, d_cur, (n ? n : 1) * sizeof(Fr));
  return s;
}

extern "C" int zolt_hyperkzg_open(zolt_ctx* ctx, const zolt_bases* params, const uint64_t* evals, size_t n, const uint64_t* point,
                                  int num_vars, uint64_t* quotients_xy, uint8_t* quotients_inf, int* n_quotients,
                                  uint64_t final_eval[4]) {
  if (!ctx || !params || !final_eval || num_vars < 0 || (n && !evals) || (num_vars && (!point || !quotients_xy || !quotients_inf)))
    return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  Fr* d_cur = nullptr;
  ZOLT_CUDA_TRY(pool_alloc(ctx, (n ? n : 1) * sizeof(Fr), (void**)&d_cur));
  if (n) ZOLT_CUDA_TRY(cudaMemcpyAsync(d_cur, evals, n * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
  const int s = hyperkzg_open_core(ctx, params, d_cur, n, point, num_vars, quotients_xy, quotients_inf, n_quotients, final_eval);
  pool_release(ctx, d_cur, (n ? n : 1) * sizeof(Fr));
  return s;
}

// ---- Dory, G1 side, and the wire formats either side of the MSM -------------------------------------------
// (pairings, G2 and GT stay on the CPU: out of scope)

// G1 points as they come off the wire -> resident bases (affine, Montgomery) + infinity flags + validity.
//   ZOLT_G1_WIRE_BE      x || y, 32-byte big-endian standard integers: raw-binary SRS (src/poly/commitment/srs.zig:65-99, :256-306);
//                        all-zero = identity; the point must satisfy y^2 = x^3 + 3 (PointNotOnCurve otherwise)
//   ZOLT_G1_WIRE_LE      the same, little-endian: the tauG1 section of a snarkjs PTAU file as the reference reads it (:616-660, :733-900)
//   ZOLT_G1_WIRE_ARK_LE  little-endian with arkworks' two flag bits in the top of y cleared, no curve check: the Jolt Dory SRS
//                        file (src/poly/commitment/dory.zig:740-751, :828-880)
__global__ void __launch_bounds__(256)
g1_from_wire_kernel(const uint8_t* __restrict__ bytes, size_t n, int format, AffineG1* __restrict__ out_xy, uint8_t* __restrict__ out_inf,
                    unsigned long long* __restrict__ first_bad) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint8_t* p = bytes + 64 * i;
    Fq raw[2];
    bool all_zero = true;
    for (int c = 0; c < 2; c++) {
      for (int w = 0; w < 8; w++) {
        uint32_t v = 0;
        for (int b = 0; b < 4; b++) {
          const int k = 4 * w + b;  // byte k of the little-endian value
          const uint8_t byte = format == ZOLT_G1_WIRE_BE ? p[32 * c + 31 - k] : p[32 * c + k];
          v |= (uint32_t)byte << (8 * b);
        }
        if (format == ZOLT_G1_WIRE_ARK_LE && c == 1 && w == 7) v &= 0x3fffffffu;
        raw[c].v[w] = v;
        all_zero = all_zero && v == 0;
      }
    }
    AffineG1 a;
    a.x = Fq::zero();
    a.y = Fq::zero();
    bool inf = all_zero;
    if (!inf) {
      a.x = fp_to_mont<FqParams>(raw[0]);  // fromBytesBE / toMontgomery: value * R mod q (values >= q reduce, as in the reference)
      a.y = fp_to_mont<FqParams>(raw[1]);
      if (format != ZOLT_G1_WIRE_ARK_LE) {
        const Fq three = fp_from_u32<FqParams>(3u);
        const Fq lhs = fq_sqr(a.y), rhs = fp_add<FqParams>(fq_mul(fq_sqr(a.x), a.x), three);
        if (lhs != rhs) atomicMin(first_bad, (unsigned long long)i);
      }
    }
    st_fq(&out_xy[i].x, a.x);
    st_fq(&out_xy[i].y, a.y);
    out_inf[i] = inf ? 1 : 0;
  }
}

// compressG1 (src/poly/commitment/dory.zig:51-76): 32 bytes = x as a little-endian standard integer with the flags in the top
// two bits of the last byte: 0x40 point at infinity (everything else zero), 0x80 when y > -y as standard integers (:162-175)
__global__ void __launch_bounds__(256)
g1_compress_kernel(const AffineG1* __restrict__ xy, const uint8_t* __restrict__ inf, size_t n, uint8_t* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint8_t* o = out + 32 * i;
    if (inf && inf[i]) {
      for (int k = 0; k < 32; k++) o[k] = 0;
      o[31] = 0x40;
      continue;
    }
    const AffineG1 p = ld_affine(xy + i);
    const Fq xs = fp_from_mont<FqParams>(p.x);
    const Fq ys = fp_from_mont<FqParams>(p.y), nys = fp_from_mont<FqParams>(fp_neg<FqParams>(p.y));
    bool positive = true;  // equal counts as positive
    for (int w = 7; w >= 0; w--) {
      if (ys.v[w] != nys.v[w]) {
        positive = ys.v[w] < nys.v[w];
        break;
      }
    }
    for (int k = 0; k < 32; k++) o[k] = (uint8_t)(xs.v[k >> 2] >> (8 * (k & 3)));
    o[31] = (uint8_t)((o[31] & 0x3f) | (positive ? 0x00 : 0x80));
  }
}

// v[i] += s * g[i]  (the first-challenge update of Dory's reduce round, dory.zig:1277-1281: scalarMul + affine add + toAffine
// per element as written; here one thread per element, double-and-add over the canonical scalar, one inversion)
__global__ void __launch_bounds__(128)
g1_scale_add_kernel(AffineG1* __restrict__ v_xy, uint8_t* __restrict__ v_inf, const AffineG1* __restrict__ g_xy,
                    const uint8_t* __restrict__ g_inf, const __grid_constant__ Fr s_std, size_t n) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  XYZZ acc = XYZZ::identity();
  if (!(g_inf && g_inf[i])) {
    const AffineG1 g = ld_affine(g_xy + i);
    bool started = false;
    for (int w = 7; w >= 0; w--) {
      for (int b = 31; b >= 0; b--) {
        if (started) acc = xyzz_dbl(acc);
        if ((s_std.v[w] >> b) & 1u) {
          xyzz_madd(acc, g);
          started = true;
        }
      }
    }
  }
  if (!v_inf[i]) xyzz_madd(acc, ld_affine(v_xy + i));
  AffineG1 a;
  const bool finite = xyzz_to_affine(acc, a);
  st_fq(&v_xy[i].x, a.x);
  st_fq(&v_xy[i].y, a.y);
  v_inf[i] = finite ? 0 : 1;
}

// ---- HyperKZG.batchOpen (src/poly/commitment/mod.zig:607-732) ------------------------------------------------------------
// evaluateMultilinear (:788-817) as written: index bit j <-> point[j]; the direct sum only for at most 1024 entries and 10
// variables -- one block, one term per thread.
__global__ void __launch_bounds__(1024) hk_mle_small_kernel(const Fr* __restrict__ evals, size_t n, const Fr* __restrict__ point,
                                                            int nv, Fr* __restrict__ out) {
  __shared__ Fr red[1024];
  const unsigned tid = threadIdx.x;
  Fr term = Fr::zero();
  if (tid < n) {
    term = evals[tid];
    unsigned bits = tid;
    for (int j = 0; j < nv; j++, bits >>= 1) {
      const Fr r = point[j];
      term = fp_mul<FrParams>(term, (bits & 1u) ? r : fp_sub<FrParams>(Fr::one(), r));
    }
  }
  red[tid] = term;
  __syncthreads();
  for (unsigned s = 512; s > 0; s >>= 1) {
    if (tid < s) red[tid] = fp_add<FrParams>(red[tid], red[tid + s]);
    __syncthreads();
  }
  if (tid == 0) *out = red[0];
}

// polys: k tables (host pointers when `on_host`, device pointers otherwise), lens[i] entries each.
static int hyperkzg_batch_open_impl(zolt_ctx* ctx, const zolt_bases* params, const void* const* polys, const size_t* lens, size_t k,
                                    bool on_host, const uint64_t* point, int num_vars, uint64_t* evaluations, uint64_t* quotients_xy,
                                    uint8_t* quotients_inf, int* n_quotients, uint64_t final_eval[4], uint64_t batching_challenge[4]) {
  if (!ctx || !params || !final_eval || !batching_challenge || num_vars < 0 || (num_vars && !point)) return ZOLT_ERR_ARG;
  if (k && (!polys || !lens || !evaluations)) return ZOLT_ERR_ARG;
  if (num_vars && k && (!quotients_xy || !quotients_inf)) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  if (n_quotients) *n_quotients = 0;
  const Fr one = Fr::one();
  auto store = [](uint64_t* dst, const Fr& x) {
    for (int j = 0; j < 4; j++) dst[j] = (uint64_t)x.v[2 * j] | ((uint64_t)x.v[2 * j + 1] << 32);
  };
  auto load = [](const uint64_t* src) {
    Fr x;
    for (int j = 0; j < 4; j++) {
      x.v[2 * j] = (uint32_t)src[j];
      x.v[2 * j + 1] = (uint32_t)(src[j] >> 32);
    }
    return x;
  };
  if (k == 0) {  // :613-621
    memset(final_eval, 0, 32);
    store(batching_challenge, one);
    return ZOLT_OK;
  }
  cudaStream_t st = ctx->stream;
  const size_t poly_size = lens[0];
  size_t max_len = 0;
  for (size_t i = 0; i < k; i++) max_len = lens[i] > max_len ? lens[i] : max_len;
  // device copies of the tables (host input) and of the point
  Fr *d_stage = nullptr, *d_point = nullptr, *d_eval = nullptr, *d_comb = nullptr;
  const size_t stage_bytes = (max_len ? max_len : 1) * sizeof(Fr), comb_bytes = (poly_size ? poly_size : 1) * sizeof(Fr);
  if (on_host) ZOLT_CUDA_TRY(pool_alloc(ctx, stage_bytes, (void**)&d_stage));
  ZOLT_CUDA_TRY(pool_alloc(ctx, (size_t)(num_vars ? num_vars : 1) * sizeof(Fr), (void**)&d_point));
  ZOLT_CUDA_TRY(pool_alloc(ctx, k * sizeof(Fr), (void**)&d_eval));
  ZOLT_CUDA_TRY(pool_alloc(ctx, comb_bytes, (void**)&d_comb));
  int status = ZOLT_OK;
  auto cuda_ok = [&](cudaError_t e) {
    if (e != cudaSuccess && status == ZOLT_OK) {
      fprintf(stderr, "[zolt_gpu] hyperkzg batch open: %s\n", cudaGetErrorString(e));
      status = ZOLT_ERR_CUDA;
    }
    return e == cudaSuccess;
  };
  if (num_vars) cuda_ok(cudaMemcpyAsync(d_point, point, (size_t)num_vars * sizeof(Fr), cudaMemcpyHostToDevice, st));
  cuda_ok(cudaMemsetAsync(d_comb, 0, comb_bytes, st));
  // Step 2 (:631-638): gamma = 0x9a8b7c6d * prod_i (r_i + 11), 1 if that is zero -- O(num_vars) host glue
  Fr gamma = fp_from_u32<FrParams>(0x9a8b7c6du);
  const Fr eleven = fp_from_u32<FrParams>(11u);
  for (int i = 0; i < num_vars; i++) gamma = fp_mul<FrParams>(gamma, fp_add<FrParams>(load(point + 4 * i), eleven));
  if (gamma.is_zero()) gamma = one;
  store(batching_challenge, gamma);
  // Steps 1 and 3 (:625-656): evaluations and combined = sum_i gamma^i poly_i over the first polynomial's length
  Fr gamma_power = one;
  for (size_t i = 0; i < k && status == ZOLT_OK; i++) {
    const Fr* d_poly = static_cast<const Fr*>(polys[i]);
    if (on_host) {
      if (lens[i]) cuda_ok(cudaMemcpyAsync(d_stage, polys[i], lens[i] * sizeof(Fr), cudaMemcpyHostToDevice, st));
      d_poly = d_stage;
    }
    if (lens[i] == 0) {
      cuda_ok(cudaMemsetAsync(d_eval + i, 0, sizeof(Fr), st));  // evaluateMultilinear of an empty table is zero
    } else if (num_vars == 0 || !(num_vars <= 10 && lens[i] <= 1024)) {
      // no variables, or the reference's fallback for larger tables: evals[0] as written (:815-816)
      cuda_ok(cudaMemcpyAsync(d_eval + i, d_poly, sizeof(Fr), cudaMemcpyDeviceToDevice, st));
    } else {
      hk_mle_small_kernel<<<1, 1024, 0, st>>>(d_poly, lens[i], d_point, num_vars, d_eval + i);
      ctx->launches++;
    }
    const size_t m = lens[i] < poly_size ? lens[i] : poly_size;
    if (m && status == ZOLT_OK) {
      uint64_t s64[4];
      store(s64, gamma_power);
      status = zolt_fr_vec_dev(ctx, ZOLT_VEC_AXPY, d_comb, d_poly, nullptr, nullptr, nullptr, s64, m);
    }
    gamma_power = fp_mul<FrParams>(gamma_power, gamma);
    if (on_host) cuda_ok(cudaStreamSynchronize(st));  // the staging buffer is reused by the next table
  }
  if (status == ZOLT_OK) {
    cuda_ok(cudaMemcpyAsync(evaluations, d_eval, k * sizeof(Fr), cudaMemcpyDeviceToHost, st));
    cuda_ok(cudaStreamSynchronize(st));
  }
  if (status == ZOLT_OK) {
    if (num_vars == 0) {
      // Step 4/5 (:658-673): nothing to fold, final_eval = sum_i gamma^i * evaluation_i
      Fr acc = Fr::zero(), gp = one;
      for (size_t i = 0; i < k; i++) {
        acc = fp_add<FrParams>(acc, fp_mul<FrParams>(gp, load(evaluations + 4 * i)));
        gp = fp_mul<FrParams>(gp, gamma);
      }
      store(final_eval, acc);
    } else {
      // Step 5 (:675-718): the same fold-and-commit loop as open(), on the combined table (consumed in place)
      status = hyperkzg_open_core(ctx, params, d_comb, poly_size, point, num_vars, quotients_xy, quotients_inf, n_quotients, final_eval);
    }
  }
  if (d_stage) pool_release(ctx, d_stage, stage_bytes);
  pool_release(ctx, d_point, (size_t)(num_vars ? num_vars : 1) * sizeof(Fr));
  pool_release(ctx, d_eval, k * sizeof(Fr));
  pool_release(ctx, d_comb, comb_bytes);
  return status;
}

extern "C" int zolt_msm_bases_upload_wire(zolt_ctx* ctx, const uint8_t* bytes, size_t n, int format, zolt_bases** out, size_t* bad_index) {
  if (!ctx || !out || (n && !bytes) || format < ZOLT_G1_WIRE_BE || format > ZOLT_G1_WIRE_ARK_LE) return ZOLT_ERR_ARG;
  *out = nullptr;
  if (bad_index) *bad_index = (size_t)-1;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  zolt_bases* b = new (std::nothrow) zolt_bases();
  if (!b) return ZOLT_ERR_ARG;
  b->ctx = ctx;
  b->n = n;
  uint8_t* d_bytes = nullptr;
  unsigned long long* d_bad = nullptr;
  int status = ZOLT_OK;
  auto ok = [&](cudaError_t e) {
    if (e != cudaSuccess && status == ZOLT_OK) {
      fprintf(stderr, "[zolt_gpu] bases from wire format: %s\n", cudaGetErrorString(e));
      status = ZOLT_ERR_CUDA;
    }
    return e == cudaSuccess;
  };
  unsigned long long bad = ~0ull;
  if (ok(cudaMalloc((void**)&b->d_xy, (n ? n : 1) * sizeof(AffineG1))) && ok(cudaMalloc((void**)&b->d_inf, n ? n : 1)) &&
      ok(cudaMalloc((void**)&d_bytes, n ? 64 * n : 1)) && ok(cudaMalloc((void**)&d_bad, 8))) {
    ok(cudaMemcpyAsync(d_bad, &bad, 8, cudaMemcpyHostToDevice, ctx->stream));
    if (n) {
      ok(cudaMemcpyAsync(d_bytes, bytes, 64 * n, cudaMemcpyHostToDevice, ctx->stream));
      size_t blocks = (n + 255) / 256;
      if (blocks > (size_t)ctx->num_sms * 8) blocks = (size_t)ctx->num_sms * 8;
      g1_from_wire_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(d_bytes, n, format, b->d_xy, b->d_inf, d_bad);
      ctx->launches++;
      ok(cudaGetLastError());
    }
    ok(cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, ctx->stream));
    ok(cudaStreamSynchronize(ctx->stream));
  }
  if (d_bytes) cudaFree(d_bytes);
  if (d_bad) cudaFree(d_bad);
  if (status == ZOLT_OK && bad != ~0ull) {  // SRSError.PointNotOnCurve
    if (bad_index) *bad_index = (size_t)bad;
    status = ZOLT_ERR_ARG;
  }
  if (status != ZOLT_OK) {
    zolt_msm_bases_free(b);
    return status;
  }
  *out = b;
  return ZOLT_OK;
}

extern "C" int zolt_msm_bases_wrap_dev(zolt_ctx* ctx, const void* d_xy, const uint8_t* d_inf, size_t n, zolt_bases** out) {
  if (!ctx || !out || (n && !d_xy)) return ZOLT_ERR_ARG;
  zolt_bases* b = new (std::nothrow) zolt_bases();
  if (!b) return ZOLT_ERR_ARG;
  b->ctx = ctx;
  b->n = n;
  b->d_xy = const_cast<AffineG1*>(static_cast<const AffineG1*>(d_xy));
  b->d_inf = const_cast<uint8_t*>(d_inf);
  b->borrowed = true;
  *out = b;
  return ZOLT_OK;
}

extern "C" int zolt_msm_bases_dev_ptrs(const zolt_bases* b, const void** d_xy, const uint8_t** d_inf) {
  if (!b || !d_xy || !d_inf) return ZOLT_ERR_ARG;
  *d_xy = b->d_xy;
  *d_inf = b->d_inf;
  return ZOLT_OK;
}

extern "C" int zolt_g1_compress(zolt_ctx* ctx, const zolt_bases* pts, size_t offset, size_t n, uint8_t* out32) {
  if (!ctx || !pts || (n && !out32) || offset > pts->n || n > pts->n - offset) return ZOLT_ERR_ARG;
  if (n == 0) return ZOLT_OK;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  uint8_t* d_out = nullptr;
  ZOLT_CUDA_TRY(pool_alloc(ctx, 32 * n, (void**)&d_out));
  size_t blocks = (n + 255) / 256;
  if (blocks > (size_t)ctx->num_sms * 8) blocks = (size_t)ctx->num_sms * 8;
  g1_compress_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(pts->d_xy + offset, pts->d_inf ? pts->d_inf + offset : nullptr, n, d_out);
  ctx->launches++;
  int status = ZOLT_OK;
  if (cudaGetLastError() != cudaSuccess || cudaMemcpyAsync(out32, d_out, 32 * n, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
      cudaStreamSynchronize(ctx->stream) != cudaSuccess)
    status = ZOLT_ERR_CUDA;
  pool_release(ctx, d_out, 32 * n);
  return status;
}

extern "C" int zolt_g1_scale_add(zolt_ctx* ctx, zolt_bases* v, const zolt_bases* g, size_t n, const uint64_t scalar[4]) {
  if (!ctx || !v || !g || !scalar || n > v->n || n > g->n) return ZOLT_ERR_ARG;
  if (n == 0) return ZOLT_OK;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  if (!v->d_inf) {  // the sums may hit the identity: the vector needs flags from now on
    if (v->borrowed) return ZOLT_ERR_ARG;
    ZOLT_CUDA_TRY(cudaMalloc((void**)&v->d_inf, v->n ? v->n : 1));
    ZOLT_CUDA_TRY(cudaMemsetAsync(v->d_inf, 0, v->n ? v->n : 1, ctx->stream));
  }
  Fr s;
  for (int j = 0; j < 4; j++) {
    s.v[2 * j] = (uint32_t)scalar[j];
    s.v[2 * j + 1] = (uint32_t)(scalar[j] >> 32);
  }
  const Fr s_std = fp_from_mont<FrParams>(s);
  g1_scale_add_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(v->d_xy, v->d_inf, g->d_xy, g->d_inf, s_std, n);
  ctx->launches++;
  ZOLT_CUDA_TRY(cudaGetLastError());
  if (v->d_pre) {  // a window-level table of the old points is stale
    ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    cudaFree(v->d_pre);
    v->d_pre = nullptr;
  }
  return ZOLT_OK;
}

// DoryCommitmentScheme.commit's G1 half (dory.zig:989-1042): sigma = (num_vars + 1) / 2, nu = num_vars - sigma, row r =
// MSM(g1_vec[0 .. 2^sigma], evals[r 2^sigma ..]) for the 2^nu rows -- one batched MSM over the shared bases.
static int dory_rows_impl(zolt_ctx* ctx, const zolt_bases* g1, const void* evals, bool on_host, size_t len, uint64_t* rows_xy,
                          uint8_t* rows_inf, size_t* num_rows, size_t* num_cols) {
  if (!ctx || !g1 || !num_rows || !num_cols || (len && (!evals || !rows_xy || !rows_inf))) return ZOLT_ERR_ARG;
  *num_rows = *num_cols = 0;
  if (len == 0) return ZOLT_OK;  // commit() returns GT.one(): no rows
  size_t num_vars = 0;
  while ((size_t(2) << num_vars) <= len) num_vars++;  // log2_int(len)
  if (len <= 1) num_vars = 1;
  const size_t sigma = (num_vars + 1) / 2, nu = num_vars - sigma;
  size_t cols = size_t(1) << sigma, rows = size_t(1) << nu;
  if (len < cols) {  // a single entry: one row with what there is (row_evals = evals[0 .. min(cols, len)])
    cols = len;
    rows = 1;
  }
  if (cols > g1->n) return ZOLT_ERR_ARG;
  *num_rows = rows;
  *num_cols = cols;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  if (on_host) return zolt_msm_batch(ctx, g1, static_cast<const uint64_t*>(evals), rows, cols, rows_xy, rows_inf);
  XYZZ* d_res = nullptr;
  ZOLT_CUDA_TRY(pool_alloc(ctx, rows * sizeof(XYZZ), (void**)&d_res));
  int status = msm_core(ctx, g1, 0, static_cast<const Fr*>(evals), rows, cols, d_res);
  if (status == ZOLT_OK) status = finish_affine(ctx, d_res, rows, rows_xy, rows_inf);
  pool_release(ctx, d_res, rows * sizeof(XYZZ));
  return status;
}

extern "C" int zolt_dory_row_commitments(zolt_ctx* ctx, const zolt_bases* g1, const uint64_t* evals, size_t len, uint64_t* rows_xy,
                                         uint8_t* rows_inf, size_t* num_rows, size_t* num_cols) {
  return dory_rows_impl(ctx, g1, evals, true, len, rows_xy, rows_inf, num_rows, num_cols);
}
extern "C" int zolt_dory_row_commitments_dev(zolt_ctx* ctx, const zolt_bases* g1, const void* d_evals, size_t len, uint64_t* rows_xy,
                                             uint8_t* rows_inf, size_t* num_rows, size_t* num_cols) {
  return dory_rows_impl(ctx, g1, d_evals, false, len, rows_xy, rows_inf, num_rows, num_cols);
}

extern "C" int zolt_hyperkzg_batch_open(zolt_ctx* ctx, const zolt_bases* params, const uint64_t* const* polys, const size_t* lens, size_t k,
                                        const uint64_t* point, int num_vars, uint64_t* evaluations, uint64_t* quotients_xy,
                                        uint8_t* quotients_inf, int* n_quotients, uint64_t final_eval[4], uint64_t batching_challenge[4]) {
  return hyperkzg_batch_open_impl(ctx, params, reinterpret_cast<const void* const*>(polys), lens, k, true, point, num_vars, evaluations,
                                  quotients_xy, quotients_inf, n_quotients, final_eval, batching_challenge);
}

extern "C" int zolt_hyperkzg_batch_open_dev(zolt_ctx* ctx, const zolt_bases* params, const void* const* d_polys, const size_t* lens, size_t k,
                                            const uint64_t* point, int num_vars, uint64_t* evaluations, uint64_t* quotients_xy,
                                            uint8_t* quotients_inf, int* n_quotients, uint64_t final_eval[4],
                                            uint64_t batching_challenge[4]) {
  return hyperkzg_batch_open_impl(ctx, params, d_polys, lens, k, false, point, num_vars, evaluations, quotients_xy, quotients_inf,
                                  n_quotients, final_eval, batching_challenge);
}

extern "C" int zolt_hyperkzg_setup(zolt_ctx* ctx, size_t max_degree, zolt_bases** out) {
  if (!ctx || !out) return ZOLT_ERR_ARG;
  *out = nullptr;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  Fr* d_s = nullptr;
  ZOLT_CUDA_TRY(cudaMalloc((void**)&d_s, (max_degree ? max_degree : 1) * sizeof(Fr)));
  if (max_degree) {
    size_t blocks = (max_degree + 255) / 256;
    if (blocks > (size_t)ctx->num_sms * 16) blocks = (size_t)ctx->num_sms * 16;
    tau_powers_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(d_s, max_degree, 0x12345678u);
    ctx->launches++;
  }
  int s = fixed_base_dev(ctx, d_s, max_degree, out);
  cudaFree(d_s);
  return s;
}
