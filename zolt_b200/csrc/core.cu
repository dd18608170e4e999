// Context, device-memory passthrough, element-wise field ops and integer-pipe micro-benchmarks.
#include <cstring>
#include <new>
#include <type_traits>

#include "common.cuh"
#include "field.cuh"

using namespace zolt;

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct zolt_timer_state {
  cudaEvent_t a = nullptr, b = nullptr;
};
static thread_local zolt_timer_state g_timer;

extern "C" const char* zolt_gpu_version(void) { return "zolt_b200 0.1 (sm_100a)"; }

extern "C" int zolt_gpu_init(int device, void* stream, zolt_ctx** out) {
  if (!out) return ZOLT_ERR_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    fprintf(stderr, "[zolt_gpu] no CUDA device (%s); this library has no CPU path\n",
            e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return ZOLT_ERR_NOGPU;
  }
  if (device < 0 || device >= count) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(device));
  zolt_ctx* ctx = new (std::nothrow) zolt_ctx();
  if (!ctx) return ZOLT_ERR_ARG;
  ctx->device = device;
  if (stream) {
    ctx->stream = (cudaStream_t)stream;
    ctx->owns_stream = false;
  } else {
    ZOLT_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->owns_stream = true;
  }
  cudaDeviceProp prop;
  ZOLT_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  ctx->num_sms = prop.multiProcessorCount;
  ZOLT_CUDA_TRY(cudaHostAlloc((void**)&ctx->h_mailbox, 64 * sizeof(uint64_t), cudaHostAllocMapped));
  ZOLT_CUDA_TRY(cudaHostGetDevicePointer((void**)&ctx->d_mailbox, ctx->h_mailbox, 0));
  memset(ctx->h_mailbox, 0, 64 * sizeof(uint64_t));
  *out = ctx;
  return ZOLT_OK;
}

extern "C" int zolt_gpu_shutdown(zolt_ctx* ctx) {
  if (!ctx) return ZOLT_ERR_ARG;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->h_mailbox) cudaFreeHost(ctx->h_mailbox);
  if (ctx->ws) cudaFree(ctx->ws);
  if (ctx->ws_in) cudaFree(ctx->ws_in);
  pool_destroy(ctx);
  for (int i = 0; i < 2 * zolt_ctx::kMaxProf; i++)
    if (ctx->prof_ev[i]) cudaEventDestroy(ctx->prof_ev[i]);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->ev_gate) cudaEventDestroy(ctx->ev_gate);
  if (ctx->ev_copied) cudaEventDestroy(ctx->ev_copied);
  if (ctx->ev_copied2) cudaEventDestroy(ctx->ev_copied2);
  if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return ZOLT_OK;
}

extern "C" int zolt_gpu_sync(zolt_ctx* ctx) {
  if (!ctx) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return ZOLT_OK;
}

extern "C" int zolt_gpu_launch_count(zolt_ctx* ctx, unsigned long long* out) {
  if (!ctx || !out) return ZOLT_ERR_ARG;
  *out = ctx->launches;
  return ZOLT_OK;
}

extern "C" int zolt_gpu_set_tuning(zolt_ctx* ctx, int key, long long value) {
  if (!ctx || value < 0) return ZOLT_ERR_ARG;
  switch (key) {
    case ZOLT_TUNE_SC_SMALL_PAIRS: ctx->sc_small_pairs = value > (1 << 14) ? (1 << 14) : value; return ZOLT_OK;
    case ZOLT_TUNE_SC_TAIL_PAIRS: ctx->sc_tail_pairs = value > (1 << 14) ? (1 << 14) : value; return ZOLT_OK;
    case ZOLT_TUNE_SC_TAIL_TIMEOUT_MS: ctx->sc_tail_timeout_ms = value; return ZOLT_OK;
    case 9: ctx->msm_sort = (int)value; return ZOLT_OK;
    case 8: ctx->msm_dbg = reinterpret_cast<unsigned long long*>(value); return ZOLT_OK;          // diagnostics: device pointer or 0
    case 7: ctx->sc_round_dbg = reinterpret_cast<unsigned long long*>(value); return ZOLT_OK;     // diagnostics: device pointer or 0
    case 6: ctx->sc_tail_dbg = reinterpret_cast<unsigned long long*>(value); return ZOLT_OK;  // diagnostics: device pointer or 0
    case ZOLT_TUNE_PEER_TIMEOUT_MS: ctx->peer_timeout_ms = value; return ZOLT_OK;
    case ZOLT_TUNE_SC_BIND_F64: ctx->sc_bind_f64 = (int)(value & 0xff); return ZOLT_OK;
    default: return ZOLT_ERR_ARG;
  }
}

extern "C" int zolt_dev_alloc(zolt_ctx* ctx, size_t bytes, void** dptr) {
  if (!ctx || !dptr) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  ZOLT_CUDA_TRY(cudaMalloc(dptr, bytes ? bytes : 1));
  return ZOLT_OK;
}

// Scratch from the context's buffer cache: what the library's own calls use for their temporaries.  A released buffer is
// handed to the next request of a similar size without cudaMalloc / cudaFree (each costs milliseconds for GiB-sized
// tables and cudaFree synchronises the device); work on the buffer must have been issued on the context's stream.
extern "C" int zolt_dev_alloc_cached(zolt_ctx* ctx, size_t bytes, void** dptr) {
  if (!ctx || !dptr) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  ZOLT_CUDA_TRY(pool_alloc(ctx, bytes, dptr));
  return ZOLT_OK;
}
extern "C" int zolt_dev_release_cached(zolt_ctx* ctx, void* dptr, size_t bytes) {
  if (!ctx) return ZOLT_ERR_ARG;
  pool_release(ctx, dptr, bytes);
  return ZOLT_OK;
}

extern "C" int zolt_dev_free(zolt_ctx* ctx, void* dptr) {
  if (!ctx) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  ZOLT_CUDA_TRY(cudaFree(dptr));
  return ZOLT_OK;
}

extern "C" int zolt_dev_upload(zolt_ctx* ctx, void* dptr, const void* hptr, size_t bytes) {
  if (!ctx || (bytes && (!dptr || !hptr))) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  ZOLT_CUDA_TRY(cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice, ctx->stream));
  ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return ZOLT_OK;
}

extern "C" int zolt_dev_download(zolt_ctx* ctx, void* hptr, const void* dptr, size_t bytes) {
  if (!ctx || (bytes && (!dptr || !hptr))) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  ZOLT_CUDA_TRY(cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return ZOLT_OK;
}

extern "C" int zolt_dev_copy(zolt_ctx* ctx, void* d_dst, const void* d_src, size_t bytes) {
  if (!ctx || (bytes && (!d_dst || !d_src))) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  ZOLT_CUDA_TRY(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return ZOLT_OK;
}

extern "C" int zolt_prof_enable(zolt_ctx* ctx, int on) {
  if (!ctx) return ZOLT_ERR_ARG;
  ctx->prof_on = on != 0;
  ctx->prof_used = 0;
  return ZOLT_OK;
}

extern "C" int zolt_prof_read(zolt_ctx* ctx, int* tags, float* ms, int max_entries, int* n_entries) {
  if (!ctx || !tags || !ms || !n_entries) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  int n = ctx->prof_used < max_entries ? ctx->prof_used : max_entries;
  for (int i = 0; i < n; i++) {
    tags[i] = ctx->prof_tag[i];
    ZOLT_CUDA_TRY(cudaEventElapsedTime(&ms[i], ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]));
  }
  *n_entries = n;
  ctx->prof_used = 0;
  return ZOLT_OK;
}

extern "C" int zolt_timer_start(zolt_ctx* ctx) {
  if (!ctx) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  if (!g_timer.a) {
    ZOLT_CUDA_TRY(cudaEventCreate(&g_timer.a));
    ZOLT_CUDA_TRY(cudaEventCreate(&g_timer.b));
  }
  ZOLT_CUDA_TRY(cudaEventRecord(g_timer.a, ctx->stream));
  return ZOLT_OK;
}

extern "C" int zolt_timer_stop_ms(zolt_ctx* ctx, float* ms) {
  if (!ctx || !ms || !g_timer.a) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaEventRecord(g_timer.b, ctx->stream));
  ZOLT_CUDA_TRY(cudaEventSynchronize(g_timer.b));
  ZOLT_CUDA_TRY(cudaEventElapsedTime(ms, g_timer.a, g_timer.b));
  return ZOLT_OK;
}

// ------------------------------------------------------------------------------------------------
// element-wise field ops
// ------------------------------------------------------------------------------------------------
template <class P>
__global__ void __launch_bounds__(256) field_batch_kernel(int op, const Fp<P>* __restrict__ a,
                                                          const Fp<P>* __restrict__ b,
                                                          Fp<P>* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    Fp<P> x = a[i];
    Fp<P> y = b ? b[i] : Fp<P>::zero();
    Fp<P> r;
    switch (op) {
      case ZOLT_OP_MUL: r = fp_mul<P>(x, y); break;
      case ZOLT_OP_ADD: r = fp_add<P>(x, y); break;
      case ZOLT_OP_SUB: r = fp_sub<P>(x, y); break;
      case ZOLT_OP_INV: r = fp_inv<P>(x); break;
      case ZOLT_OP_FROM_MONT: r = fp_from_mont<P>(x); break;
      case ZOLT_OP_TO_MONT: r = fp_to_mont<P>(x); break;
      default: r = fp_neg<P>(x); break;
    }
    out[i] = r;
  }
}

extern "C" int zolt_field_batch_op(zolt_ctx* ctx, int field, int op, const uint64_t* a, const uint64_t* b,
                                   uint64_t* out, size_t n) {
  if (!ctx || !a || !out) return ZOLT_ERR_ARG;
  if (op < ZOLT_OP_MUL || op > ZOLT_OP_NEG) return ZOLT_ERR_ARG;
  if (field != ZOLT_FIELD_FR && field != ZOLT_FIELD_FQ) return ZOLT_ERR_ARG;
  const bool binary = (op == ZOLT_OP_MUL || op == ZOLT_OP_ADD || op == ZOLT_OP_SUB);
  if (binary && !b) return ZOLT_ERR_ARG;
  if (n == 0) return ZOLT_OK;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  void *da = nullptr, *db = nullptr, *dout = nullptr;
  const size_t bytes = n * 32;
  ZOLT_CUDA_TRY(cudaMalloc(&da, bytes));
  ZOLT_CUDA_TRY(cudaMalloc(&dout, bytes));
  ZOLT_CUDA_TRY(cudaMemcpyAsync(da, a, bytes, cudaMemcpyHostToDevice, ctx->stream));
  if (binary) {
    ZOLT_CUDA_TRY(cudaMalloc(&db, bytes));
    ZOLT_CUDA_TRY(cudaMemcpyAsync(db, b, bytes, cudaMemcpyHostToDevice, ctx->stream));
  }
  size_t blocks = (n + 255) / 256;
  if (blocks > (size_t)ctx->num_sms * 8) blocks = (size_t)ctx->num_sms * 8;
  if (field == ZOLT_FIELD_FR)
    field_batch_kernel<FrParams><<<(unsigned)blocks, 256, 0, ctx->stream>>>(op, (const Fr*)da, (const Fr*)db, (Fr*)dout, n);
  else
    field_batch_kernel<FqParams><<<(unsigned)blocks, 256, 0, ctx->stream>>>(op, (const Fq*)da, (const Fq*)db, (Fq*)dout, n);
  ctx->launches++;
  ZOLT_CUDA_TRY(cudaGetLastError());
  ZOLT_CUDA_TRY(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  ZOLT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  cudaFree(da);
  cudaFree(dout);
  if (db) cudaFree(db);
  return ZOLT_OK;
}

// ------------------------------------------------------------------------------------------------
// integer-pipe micro-benchmarks
// ------------------------------------------------------------------------------------------------
// 8 independent dependency chains per thread so a single warp can keep the pipe busy; 256 threads x
// 8 blocks/SM resident.  Every result feeds the final store so nothing is dead code.
template <int KIND>
__global__ void __launch_bounds__(256) int_pipe_kernel(uint32_t* out, int iters, uint32_t seed) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (KIND == 0 || KIND == 4) {
    uint32_t x[8], m = seed | 1u, c = tid;
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = tid + j;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
          if (KIND == 0) {
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(m), "r"(c));
          } else {
            asm volatile("mad.lo.u32 %0, %0, 0x43e1f593, %1;" : "+r"(x[j]) : "r"(c));
          }
        }
      }
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s ^= x[j];
    out[tid] = s;
  } else if (KIND == 1) {
    unsigned long long x[8];
    uint32_t m = seed | 1u;
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = tid + j;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
          uint32_t lo = (uint32_t)x[j];
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[j]) : "r"(lo), "r"(m));
        }
      }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s ^= x[j];
    out[tid] = (uint32_t)s ^ (uint32_t)(s >> 32);
  } else if (KIND == 5 || KIND == 6) {
    // KIND 5: FP64 fused multiply-adds alone; KIND 6: the same DFMA stream interleaved 1:1 with IMADs,
    // to see whether the FP64 and the integer multiply pipes issue concurrently.
    double d[8], dm = 1.0000001 + (double)(seed & 3), dc = (double)tid;
    uint32_t x[8], m = seed | 1u, c = tid;
#pragma unroll
    for (int j = 0; j < 8; j++) { d[j] = (double)(tid + j); x[j] = tid + j; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
          asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[j]) : "d"(dm), "d"(dc));
          if (KIND == 6) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(m), "r"(c));
        }
      }
    }
    double s = 0;
    uint32_t sx = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) { s += d[j]; sx ^= x[j]; }
    out[tid] = (uint32_t)__double2ll_rn(s) ^ sx;
  } else {
    // field products: 2 independent chains
    using P = typename std::conditional<KIND == 2, FrParams, FqParams>::type;
    Fp<P> a, b, c;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      a.v[j] = P::one(j) ^ (tid & 0xff);
      b.v[j] = P::r2(j);
      c.v[j] = P::one(j) ^ ((tid >> 8) & 0xff);
    }
    a.v[7] &= 0x0fffffffu;
    c.v[7] &= 0x0fffffffu;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        a = fp_mul<P>(a, b);
        c = fp_mul<P>(c, b);
      }
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s ^= a.v[j] ^ c.v[j];
    out[tid] = s;
  }
}

extern "C" int zolt_bench_int_pipe_cfg(zolt_ctx* ctx, int kind, int iters, int threads, int blocks_per_sm, double* ops_per_s);
extern "C" int zolt_bench_int_pipe(zolt_ctx* ctx, int kind, int iters, double* ops_per_s) {
  return zolt_bench_int_pipe_cfg(ctx, kind, iters, 256, 8, ops_per_s);
}

extern "C" int zolt_bench_int_pipe_cfg(zolt_ctx* ctx, int kind, int iters, int threads, int blocks_per_sm,
                                       double* ops_per_s) {
  if (!ctx || !ops_per_s || kind < 0 || kind > 6 || iters <= 0) return ZOLT_ERR_ARG;
  if (threads < 32 || threads > 256 || (threads & 31) || blocks_per_sm < 1 || blocks_per_sm > 16) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  const int blocks = ctx->num_sms * blocks_per_sm;
  uint32_t* d_out = nullptr;
  ZOLT_CUDA_TRY(cudaMalloc(&d_out, (size_t)threads * blocks * sizeof(uint32_t)));
  cudaEvent_t e0, e1;
  ZOLT_CUDA_TRY(cudaEventCreate(&e0));
  ZOLT_CUDA_TRY(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {  // first rep is the warm-up
    ZOLT_CUDA_TRY(cudaEventRecord(e0, ctx->stream));
    switch (kind) {
      case 0: int_pipe_kernel<0><<<blocks, threads, 0, ctx->stream>>>(d_out, iters, 12345u); break;
      case 1: int_pipe_kernel<1><<<blocks, threads, 0, ctx->stream>>>(d_out, iters, 12345u); break;
      case 2: int_pipe_kernel<2><<<blocks, threads, 0, ctx->stream>>>(d_out, iters, 12345u); break;
      case 3: int_pipe_kernel<3><<<blocks, threads, 0, ctx->stream>>>(d_out, iters, 12345u); break;
      case 4: int_pipe_kernel<4><<<blocks, threads, 0, ctx->stream>>>(d_out, iters, 12345u); break;
      case 5: int_pipe_kernel<5><<<blocks, threads, 0, ctx->stream>>>(d_out, iters, 12345u); break;
      default: int_pipe_kernel<6><<<blocks, threads, 0, ctx->stream>>>(d_out, iters, 12345u); break;
    }
    ctx->launches++;
    ZOLT_CUDA_TRY(cudaEventRecord(e1, ctx->stream));
    ZOLT_CUDA_TRY(cudaEventSynchronize(e1));
    ZOLT_CUDA_TRY(cudaGetLastError());
    float ms = 0;
    ZOLT_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  const double per_thread = (kind == 2 || kind == 3) ? 8.0 * iters : 64.0 * iters;
  *ops_per_s = per_thread * threads * blocks / (best * 1e-3);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  return ZOLT_OK;
}
