// Shared host-side helpers for the C-ABI implementation files.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

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
  // counters for the bench's "gpu_launches" claim
  unsigned long long launches = 0;
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
  PROF_SC_FUSED = 17     // fused bind + next-round evaluation kernel
};
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
}  // namespace zolt
