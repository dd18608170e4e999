// NVLink peer mailbox management (CUDA IPC) and the stand-alone mailbox all-gather.  See peer.cuh.
#include <cstring>
#include <new>

#include "common.cuh"
#include "peer.cuh"

using namespace zolt;

namespace {

struct GatherSrc {
  const Fr* p[kPeerSlotFr];
};

// Publish n_fr elements (one pointer each), wait for every rank, write all rows to `out`:
// transpose == 0: out[g * n_fr + i] (rank-major), transpose == 1: out[i * world + g] (one table per element).
__global__ void peer_allgather_kernel(const __grid_constant__ GatherSrc src, int n_fr, const __grid_constant__ PeerDev peer, Fr* out,
                                      int transpose) {
  __shared__ int ok;
  if (threadIdx.x == 0) {
    Fr vals[kPeerSlotFr];
#pragma unroll 1
    for (int i = 0; i < n_fr; i++) vals[i] = src.p[i][0];
    peer_publish(peer, vals, n_fr);
    ok = peer_wait(peer) ? 1 : 0;
  }
  __syncthreads();
  if (!ok) return;
  for (int j = threadIdx.x; j < peer.world * n_fr; j += blockDim.x) {
    const int g = j / n_fr, i = j % n_fr;
    const Fr v = peer_load(peer, g, i);
    Fr* dst = out + (transpose ? (size_t)i * peer.world + g : (size_t)g * n_fr + i);
    uint4* q = reinterpret_cast<uint4*>(dst);
    q[0] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
    q[1] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
  }
}

// Every rank stores its k local tables into the tail buffer of every rank (its own included) over NVLink, then the
// usual fence + flag + wait.  Table t of the gathered set starts at entry t * world * n_local of the buffer.
__global__ void __launch_bounds__(256)
peer_gather_tables_kernel(const __grid_constant__ GatherSrc src, int k, unsigned n_local, const __grid_constant__ PeerDev peer,
                          size_t buf_offset) {
  __shared__ int ok;
  const unsigned total = (unsigned)k * n_local;
  for (unsigned idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const unsigned t = idx / n_local, i = idx - t * n_local;
    const uint4* s = reinterpret_cast<const uint4*>(src.p[t] + i);
    const uint4 a = s[0], b = s[1];
    const size_t dst_entry = (size_t)t * peer.world * n_local + (size_t)peer.rank * n_local + i;
#pragma unroll 1
    for (int g = 0; g < peer.world; g++) {
      uint4* d = reinterpret_cast<uint4*>(peer.box[g] + buf_offset) + 2 * dst_entry;
      d[0] = a;
      d[1] = b;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    peer_publish(peer, nullptr, 0);
    ok = peer_wait(peer) ? 1 : 0;
  }
  __syncthreads();
}

}  // namespace

namespace zolt {
int peer_gather_tables(zolt_peer_box* box, const Fr* const* d_src, int k, size_t n_local, const Fr** d_tables_out) {
  if (!box || !box->connected || !d_src || !d_tables_out || k < 1 || k > kPeerTailTables || n_local < 1 || n_local > kPeerTailEntries)
    return ZOLT_ERR_ARG;
  zolt_ctx* ctx = box->ctx;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  GatherSrc s{};
  for (int i = 0; i < k; i++) s.p[i] = d_src[i];
  const size_t buf_offset = kPeerHeaderBytes + (size_t)(box->gathers++ & 1u) * kPeerTailBufBytes;
  const PeerDev p = peer_next(box);
  peer_gather_tables_kernel<<<1, 256, 0, ctx->stream>>>(s, k, (unsigned)n_local, p, buf_offset);
  ctx->launches++;
  ZOLT_CUDA_TRY(cudaGetLastError());
  const Fr* base = reinterpret_cast<const Fr*>(box->d_local + buf_offset);
  for (int t = 0; t < k; t++) d_tables_out[t] = base + (size_t)t * box->world * n_local;
  return ZOLT_OK;
}
int peer_allgather_ptrs(zolt_peer_box* box, const Fr* const* d_src, int n_fr, Fr* d_out, int transpose) {
  if (!box || !box->connected || !d_src || !d_out || n_fr < 1 || n_fr > kPeerSlotFr) return ZOLT_ERR_ARG;
  zolt_ctx* ctx = box->ctx;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  GatherSrc s{};
  for (int i = 0; i < n_fr; i++) s.p[i] = d_src[i];
  const PeerDev p = peer_next(box);
  peer_allgather_kernel<<<1, 64, 0, ctx->stream>>>(s, n_fr, p, d_out, transpose);
  ctx->launches++;
  ZOLT_CUDA_TRY(cudaGetLastError());
  return ZOLT_OK;
}
int peer_check(zolt_peer_box* box) {
  if (box && box->h_err && *box->h_err) {
    fprintf(stderr, "[zolt_gpu] peer mailbox: rank %d never saw rank %d's data (timeout)\n", box->rank, *box->h_err - 1);
    return ZOLT_ERR_PEER;
  }
  return ZOLT_OK;
}
}  // namespace zolt

extern "C" int zolt_peer_box_create(zolt_ctx* ctx, zolt_peer_box** out, unsigned char handle_out[64]) {
  if (!ctx || !out || !handle_out) return ZOLT_ERR_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  static_assert(kPeerTailTables <= kPeerSlotFr, "GatherSrc holds kPeerSlotFr pointers");
  *out = nullptr;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  zolt_peer_box* box = new (std::nothrow) zolt_peer_box();
  if (!box) return ZOLT_ERR_ARG;
  box->ctx = ctx;
  // plain cudaMalloc (not a pool): IPC export needs a whole allocation
  ZOLT_CUDA_TRY(cudaMalloc((void**)&box->d_local, kPeerBoxBytes));
  ZOLT_CUDA_TRY(cudaMemset(box->d_local, 0, kPeerBoxBytes));
  ZOLT_CUDA_TRY(cudaHostAlloc((void**)&box->h_err, 64, cudaHostAllocMapped));
  ZOLT_CUDA_TRY(cudaHostGetDevicePointer((void**)&box->d_err, box->h_err, 0));
  *box->h_err = 0;
  cudaIpcMemHandle_t h;
  ZOLT_CUDA_TRY(cudaIpcGetMemHandle(&h, box->d_local));
  memcpy(handle_out, &h, 64);
  ZOLT_CUDA_TRY(cudaDeviceSynchronize());
  *out = box;
  return ZOLT_OK;
}

extern "C" int zolt_peer_box_connect(zolt_peer_box* box, const unsigned char* handles, int rank, int world) {
  if (!box || !handles || world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world || box->connected) return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(box->ctx->device));
  for (int g = 0; g < world; g++) {
    if (g == rank) {
      box->mapped[g] = box->d_local;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * (size_t)g, 64);
    void* p = nullptr;
    ZOLT_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    box->mapped[g] = (unsigned char*)p;
    box->opened[g] = true;
  }
  box->rank = rank;
  box->world = world;
  box->seq = 0;
  box->connected = true;
  return ZOLT_OK;
}

// Ranks that live in ONE process (several contexts, e.g. a multi-GPU host thread pool or the single-GPU test of the
// protocol): the mailboxes are addressed directly, no IPC handles.  Different devices need peer access.
extern "C" int zolt_peer_box_connect_local(zolt_peer_box* box, zolt_peer_box* const* boxes, int rank, int world) {
  if (!box || !boxes || world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world || box->connected || boxes[rank] != box)
    return ZOLT_ERR_ARG;
  ZOLT_CUDA_TRY(cudaSetDevice(box->ctx->device));
  for (int g = 0; g < world; g++) {
    if (!boxes[g] || !boxes[g]->d_local) return ZOLT_ERR_ARG;
    if (boxes[g]->ctx->device != box->ctx->device) {
      const cudaError_t e = cudaDeviceEnablePeerAccess(boxes[g]->ctx->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ZOLT_CUDA_TRY(e);
      (void)cudaGetLastError();
    }
    box->mapped[g] = boxes[g]->d_local;
  }
  box->rank = rank;
  box->world = world;
  box->seq = 0;
  box->connected = true;
  return ZOLT_OK;
}

extern "C" int zolt_peer_box_free(zolt_peer_box* box) {
  if (!box) return ZOLT_ERR_ARG;
  cudaSetDevice(box->ctx->device);
  cudaStreamSynchronize(box->ctx->stream);
  for (int g = 0; g < kPeerMaxWorld; g++)
    if (box->opened[g]) cudaIpcCloseMemHandle(box->mapped[g]);
  if (box->d_local) cudaFree(box->d_local);
  if (box->h_err) cudaFreeHost(box->h_err);
  delete box;
  return ZOLT_OK;
}

// Stream-ordered barrier across the ranks: an empty exchange.  Returns at once; work queued behind it on this
// rank's stream starts only after every rank's stream has reached its own barrier.
extern "C" int zolt_peer_barrier(zolt_peer_box* box) {
  if (!box || !box->connected) return ZOLT_ERR_ARG;
  zolt_ctx* ctx = box->ctx;
  ZOLT_CUDA_TRY(cudaSetDevice(ctx->device));
  GatherSrc s{};
  peer_allgather_kernel<<<1, 32, 0, ctx->stream>>>(s, 0, peer_next(box), nullptr, 0);
  ctx->launches++;
  ZOLT_CUDA_TRY(cudaGetLastError());
  return ZOLT_OK;
}

extern "C" int zolt_peer_allgather(zolt_peer_box* box, const void* d_src, int n_fr, void* d_out, int transpose) {
  if (!box || !d_src || n_fr < 1 || n_fr > kPeerSlotFr) return ZOLT_ERR_ARG;
  const Fr* ptrs[kPeerSlotFr];
  for (int i = 0; i < n_fr; i++) ptrs[i] = (const Fr*)d_src + i;
  ZOLT_TRY(peer_allgather_ptrs(box, ptrs, n_fr, (Fr*)d_out, transpose));
  ZOLT_CUDA_TRY(cudaStreamSynchronize(box->ctx->stream));
  return peer_check(box);
}
