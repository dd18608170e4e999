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
  // counters for the bench's "gpu_launches" claim
  unsigned long long launches = 0;
};
