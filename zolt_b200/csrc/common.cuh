// Shared host-side helpers for the C-ABI implementation files.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <utility>
#include <vector>

#include "../../include/zolt_gpu.h"

namespace zolt {

constexpr int kNumSMsB200 = 148;

// Every CUDA failure is reported (stderr) and surfaced as a non-zero status: there is no CPU path
// to fall back to (the Zig shim panics).
#define ZOLT_CUDA_TRY(expr)                                                                      \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      fprintf(stderr, "[zolt_gpu] %s:%d: %s -> %s\n", __FILE__, __LINE__, #expr,                 \
              cudaGetErrorString(_e));                                                           \
      return ZOLT_ERR_CUDA;                                                                      \
    }                                                                                            \
  } while (0)

// Runs `f` when the scope is left unless dismiss() was called: the C-ABI entry points return early through ZOLT_CUDA_TRY /
// ZOLT_TRY, this releases what they had built so far.
template <class F>
struct OnFail {
  F f;
  bool armed = true;
  explicit OnFail(F fn) : f(fn) {}
  OnFail(const OnFail&) = delete;
  OnFail& operator=(const OnFail&) = delete;
  ~OnFail() {
    if (armed) f();
  }
  void dismiss() { armed = false; }
};

#define ZOLT_TRY(expr)                 \
  do {                                 \
    int _s = (expr);                   \
    if (_s != ZOLT_OK) return _s;      \
  } while (0)

}  // namespace zolt

// Opaque context: one device, one stream (owned or borrowed), scratch buffers.
struct zolt_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool owns_stream = false;
  int num_sms = zolt::kNumSMsB200;
  // pinned host mailbox for small results (round evaluations, MSM result)
  uint64_t* h_mailbox = nullptr;  // 64 x u64
  uint64_t* d_mailbox = nullptr;
  // grow-only device workspace reused by every MSM on this context (cudaMalloc/cudaFree of a
  // multi-GiB arena costs more than the MSM itself)
  void* ws = nullptr;
  size_t ws_bytes = 0;
  void* ws_in = nullptr;  // staging for host scalars (zolt_msm / zolt_msm_batch)
  size_t ws_in_bytes = 0;
  // recycled device buffers (sumcheck ping-pong tables, small scratch): cudaMalloc/cudaFree per
  // instance would cost more than the rounds themselves
  std::vector<std::pair<void*, size_t>> pool_free;
  size_t pool_bytes = 0;
  std::vector<void*> pinned_free;  // recycled 256-byte mapped pinned mailboxes (one per live instance)
  // counters for the bench's "gpu_launches" claim
  unsigned long long launches = 0;
  // tuning (zolt_gpu_set_tuning)
  long long sc_small_pairs = 1 << 14;
  long long sc_tail_pairs = 1 << 12;   // rounds with at most this many pairs run inside the resident tail kernel (0: never)
  unsigned long long* sc_round_dbg = nullptr;  // diagnostics: device buffer (64 x 8 words) for the round kernels' time stamps (tuning key 7)
  int msm_sort = 0;  // tuning key 9: 0 = two-pass bucket sort where it applies, 1 = always the atomic scatter
  unsigned long long* msm_dbg = nullptr;       // diagnostics: per-warp (start, end, SM) of the level-0 accumulation kernel (tuning key 8)
  unsigned long long* sc_tail_dbg = nullptr;  // diagnostics: device buffer for the tail kernel's per-round time stamps (tuning key 6)
  long long sc_tail_timeout_ms = 0;    // how long the tail kernel waits for the host's next challenge (0: 10 s)
  std::vector<cudaStream_t> tail_streams_free;  // recycled private streams of tail kernels
  long long peer_timeout_ms = 0;  // 0: default (4 s); ZOLT_TUNE_PEER_TIMEOUT_MS
  int sc_bind_f64 = 0;  // mask of the tables whose binds run on the FP64 pipe in a fused sumcheck pass (ZOLT_TUNE_SC_BIND_F64; measured: no gain)
  // second stream + events for the host-scalar MSM: the second half of the scalars crosses PCIe during the first half's MSM
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_gate = nullptr, ev_copied = nullptr, ev_copied2 = nullptr;
  unsigned mail_seq = 0;  // last sequence number handed to a kernel that reports through a mapped host mailbox
  // phase profiling (CUDA events on `stream`): begin/end pairs tagged by the caller
  static constexpr int kMaxProf = 96;
  cudaEvent_t prof_ev[2 * kMaxProf] = {};
  int prof_tag[kMaxProf] = {};
  int prof_used = 0;
  bool prof_on = false;
};

namespace zolt {
// Tags reported by zolt_prof_read
enum ProfTag {
  PROF_MSM_DIGITS = 1,
  PROF_MSM_SCAN = 2,
  PROF_MSM_SCATTER = 3,
  PROF_MSM_ACCUM0 = 4,   // level-0 bucket accumulation (mixed additions): the dominant MSM kernel
  PROF_MSM_EDGES = 5,
  PROF_MSM_REDUCE = 6,
  PROF_MSM_AFFINE = 7,
  PROF_SC_EVAL = 16,     // round-0 evaluation kernel
  PROF_SC_FUSED = 17,    // fused bind + next-round evaluation kernel
  PROF_SC_TAIL = 18      // resident tail kernel: all rounds below the tail threshold, host round trips included
};
// Stream-ordered reuse: buffers are only handed out again to work queued on the same stream.
inline cudaError_t pool_alloc(zolt_ctx* ctx, size_t bytes, void** out) {
  if (bytes == 0) bytes = 32;
  int best = -1;
  for (int i = 0; i < (int)ctx->pool_free.size(); i++) {
    const size_t cap = ctx->pool_free[i].second;
    if (cap >= bytes && cap <= 2 * bytes + 4096 && (best < 0 || cap < ctx->pool_free[best].second)) best = i;
  }
  if (best >= 0) {
    *out = ctx->pool_free[best].first;
    ctx->pool_free.erase(ctx->pool_free.begin() + best);
    return cudaSuccess;
  }
  cudaError_t e = cudaMalloc(out, bytes);
  if (e == cudaSuccess) ctx->pool_bytes += bytes;
  return e;
}
inline size_t pool_capacity_of(size_t bytes) { return bytes == 0 ? 32 : bytes; }
inline void pool_release(zolt_ctx* ctx, void* p, size_t bytes) {
  if (p) ctx->pool_free.emplace_back(p, pool_capacity_of(bytes));
}
inline void pool_destroy(zolt_ctx* ctx) {
  for (cudaStream_t st : ctx->tail_streams_free) cudaStreamDestroy(st);
  ctx->tail_streams_free.clear();
  for (auto& b : ctx->pool_free) cudaFree(b.first);
  ctx->pool_free.clear();
  for (void* p : ctx->pinned_free) cudaFreeHost(p);
  ctx->pinned_free.clear();
}
inline cudaError_t pinned_mailbox_alloc(zolt_ctx* ctx, void** host, void** dev) {
  if (!ctx->pinned_free.empty()) {
    *host = ctx->pinned_free.back();
    ctx->pinned_free.pop_back();
  } else {
    cudaError_t e = cudaHostAlloc(host, 256, cudaHostAllocMapped);
    if (e != cudaSuccess) return e;
  }
  return cudaHostGetDevicePointer(dev, *host, 0);
}
inline void pinned_mailbox_release(zolt_ctx* ctx, void* host) {
  if (host) ctx->pinned_free.push_back(host);
}

inline int prof_begin(zolt_ctx* ctx, int tag) {
  if (!ctx->prof_on || ctx->prof_used >= zolt_ctx::kMaxProf) return -1;
  const int slot = ctx->prof_used++;
  ctx->prof_tag[slot] = tag;
  if (!ctx->prof_ev[2 * slot]) {
    cudaEventCreate(&ctx->prof_ev[2 * slot]);
    cudaEventCreate(&ctx->prof_ev[2 * slot + 1]);
  }
  cudaEventRecord(ctx->prof_ev[2 * slot], ctx->stream);
  return slot;
}
inline void prof_end(zolt_ctx* ctx, int slot) {
  if (slot >= 0) cudaEventRecord(ctx->prof_ev[2 * slot + 1], ctx->stream);
}
// the same on another stream of the context (the tail kernel's private stream)
inline int prof_begin_on(zolt_ctx* ctx, int tag, cudaStream_t stream) {
  if (!ctx->prof_on || ctx->prof_used >= zolt_ctx::kMaxProf) return -1;
  const int slot = ctx->prof_used++;
  ctx->prof_tag[slot] = tag;
  if (!ctx->prof_ev[2 * slot]) {
    cudaEventCreate(&ctx->prof_ev[2 * slot]);
    cudaEventCreate(&ctx->prof_ev[2 * slot + 1]);
  }
  cudaEventRecord(ctx->prof_ev[2 * slot], stream);
  return slot;
}
inline void prof_end_on(zolt_ctx* ctx, int slot, cudaStream_t stream) {
  if (slot >= 0) cudaEventRecord(ctx->prof_ev[2 * slot + 1], stream);
}
}  // namespace zolt
