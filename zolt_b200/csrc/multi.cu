// One process, several devices: the entry point `zolt prove` would use (a single process, src/zkvm/mod.zig:366-462).
//
// SURVEY 8b proposed zolt_gpu_init(device_ids, n_devices): this is it.  A zolt_multi owns one context (device + stream) per
// entry of device_ids and one NVLink mailbox per context, connected to each other directly (same address space: no IPC
// handles, peer access enabled between distinct devices).  The per-rank handles are ordinary zolt_ctx / zolt_peer_box, so
// everything the one-process-per-GPU deployment uses (zolt_msm_sharded*, zolt_sc_set_peers, the C++ round loop
// proveSumcheckSharded) runs unchanged with one host thread per device.  zolt_multi_msm is that pattern packaged for the
// MSM: ParallelMSM.compute's point-range split (src/msm/mod.zig:588-653) with one host thread and one GPU per chunk.
// The same device may be listed several times (ranks then share it): that is how a one-GPU box tests this path.
#include <atomic>
#include <new>
#include <thread>
#include <vector>

#include "common.cuh"
#include "peer.cuh"

using namespace zolt;

struct zolt_multi {
  std::vector<zolt_ctx*> ctx;
  std::vector<zolt_peer_box*> box;
  std::vector<void*> d_part;  // per rank: 128 bytes for its partial sum
  void* d_gather = nullptr;   // on rank 0's device: n * 128 bytes
};

extern "C" int zolt_multi_shutdown(zolt_multi* m) {
  if (!m) return ZOLT_ERR_ARG;
  for (size_t r = 0; r < m->ctx.size(); r++) {
    if (r < m->d_part.size() && m->d_part[r]) zolt_dev_free(m->ctx[r], m->d_part[r]);
  }
  if (m->d_gather && !m->ctx.empty()) zolt_dev_free(m->ctx[0], m->d_gather);
  for (zolt_peer_box* b : m->box)
    if (b) zolt_peer_box_free(b);
  for (zolt_ctx* c : m->ctx)
    if (c) zolt_gpu_shutdown(c);
  delete m;
  return ZOLT_OK;
}

extern "C" int zolt_gpu_init_multi(const int* device_ids, int n_devices, zolt_multi** out) {
  if (!device_ids || !out || n_devices < 1 || n_devices > kPeerMaxWorld) return ZOLT_ERR_ARG;
  *out = nullptr;
  zolt_multi* m = new (std::nothrow) zolt_multi();
  if (!m) return ZOLT_ERR_ARG;
  int status = ZOLT_OK;
  for (int r = 0; r < n_devices && status == ZOLT_OK; r++) {
    zolt_ctx* c = nullptr;
    status = zolt_gpu_init(device_ids[r], nullptr, &c);
    m->ctx.push_back(c);
  }
  for (int r = 0; r < n_devices && status == ZOLT_OK; r++) {
    zolt_peer_box* b = nullptr;
    unsigned char handle[64];
    status = zolt_peer_box_create(m->ctx[r], &b, handle);
    m->box.push_back(b);
  }
  for (int r = 0; r < n_devices && status == ZOLT_OK; r++) status = zolt_peer_box_connect_local(m->box[r], m->box.data(), r, n_devices);
  for (int r = 0; r < n_devices && status == ZOLT_OK; r++) {
    void* p = nullptr;
    status = zolt_dev_alloc(m->ctx[r], 128, &p);
    m->d_part.push_back(p);
  }
  if (status == ZOLT_OK) status = zolt_dev_alloc(m->ctx[0], 128 * (size_t)n_devices, &m->d_gather);
  if (status != ZOLT_OK) {
    zolt_multi_shutdown(m);
    return status;
  }
  *out = m;
  return ZOLT_OK;
}

extern "C" int zolt_multi_size(const zolt_multi* m, int* n) {
  if (!m || !n) return ZOLT_ERR_ARG;
  *n = (int)m->ctx.size();
  return ZOLT_OK;
}

extern "C" int zolt_multi_rank(zolt_multi* m, int rank, zolt_ctx** ctx, zolt_peer_box** box) {
  if (!m || rank < 0 || rank >= (int)m->ctx.size()) return ZOLT_ERR_ARG;
  if (ctx) *ctx = m->ctx[rank];
  if (box) *box = m->box[rank];
  return ZOLT_OK;
}

// MSM over point-range shards: bases[r] holds rank r's contiguous range of the bases (resident on device r), `scalars` is
// the whole host vector; rank r takes the next bases[r]->n scalars.  One host thread per device computes its partial
// (upload overlapped with the multiplication, as in zolt_msm); the 128-byte partials are copied to device 0 and summed
// there.  == MSM.compute(all bases, scalars).
extern "C" int zolt_multi_msm(zolt_multi* m, zolt_bases* const* bases, const uint64_t* scalars, size_t n_total, uint64_t out_xy[8],
                              uint8_t* out_inf) {
  if (!m || !bases || (n_total && !scalars) || !out_xy || !out_inf) return ZOLT_ERR_ARG;
  const int world = (int)m->ctx.size();
  std::vector<size_t> first((size_t)world + 1, 0);
  for (int r = 0; r < world; r++) {
    size_t n = 0;
    if (!bases[r] || zolt_msm_bases_len(bases[r], &n) != ZOLT_OK) return ZOLT_ERR_ARG;
    first[(size_t)r + 1] = first[(size_t)r] + n;
  }
  if (first[(size_t)world] != n_total) return ZOLT_ERR_ARG;
  std::vector<int> status((size_t)world, ZOLT_OK);
  std::vector<std::thread> pool;
  for (int r = 0; r < world; r++) {
    pool.emplace_back([&, r]() {
      zolt_ctx* c = m->ctx[(size_t)r];
      const size_t n = first[(size_t)r + 1] - first[(size_t)r];
      int st = zolt_msm_partial(c, bases[r], 0, scalars + 4 * first[(size_t)r], n, m->d_part[(size_t)r]);
      if (st == ZOLT_OK) st = zolt_gpu_sync(c);
      status[(size_t)r] = st;
    });
  }
  for (auto& t : pool) t.join();
  for (int r = 0; r < world; r++)
    if (status[(size_t)r] != ZOLT_OK) return status[(size_t)r];
  zolt_ctx* c0 = m->ctx[0];
  ZOLT_CUDA_TRY(cudaSetDevice(c0->device));
  for (int r = 0; r < world; r++)
    ZOLT_CUDA_TRY(cudaMemcpyPeerAsync(static_cast<char*>(m->d_gather) + 128 * (size_t)r, c0->device, m->d_part[(size_t)r],
                                      m->ctx[(size_t)r]->device, 128, c0->stream));
  return zolt_g1_sum_partials(c0, m->d_gather, (size_t)world, out_xy, out_inf);
}
