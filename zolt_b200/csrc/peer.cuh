// NVLink peer mailbox: the exchange step of the sharded hot path, done by the producing kernel itself.
//
// across GPUs the sumcheck exchanges (d+1) field elements per round and the MSM one 128-byte
// partial per call.  An NCCL all-gather for 128 bytes is all launch latency (one more kernel, host-side
// enqueue, ~100 us per round through the harness: profiles/r1_v4_bench_g2.json), so the few bytes travel
// differently here: every rank owns a small mailbox in its HBM, exported with CUDA IPC and mapped by every
// other rank (one process per GPU, NVSwitch makes every peer one hop away).  The thread that holds a kernel's
// result STORES it into every rank's mailbox over NVLink, fences, and raises a per-(slot, sender) flag; it then
// waits until the flags of all ranks have reached the same exchange number in its OWN mailbox (a local spin)
// and folds the rows.  No second kernel, no host involvement, no collective library on the data path.
//
// Exchanges are numbered 1, 2, 3, ... identically on every rank (same call sequence, one stream per rank).
// Exchange s uses slot s & 1.  Two slots are enough: a rank can only publish exchange s+2 after it has
// completed s+1, which needs every peer's s+1 publish, which each peer issues (stream order) after its own
// kernel of exchange s has read slot s & 1.
#pragma once
#include "common.cuh"
#include "field.cuh"

namespace zolt {

constexpr int kPeerMaxWorld = 8;  // one NVSwitch node of B200s
constexpr int kPeerSlotFr = 8;    // 256 bytes per (slot, sender)
constexpr size_t kPeerDataBytes = 2ull * kPeerMaxWorld * kPeerSlotFr * 32;
constexpr size_t kPeerHeaderBytes = 8192;  // rows + flags, padded
// Table-gather area (the tail of a sharded sumcheck): two buffers of kPeerTailTables tables, each table holding
// kPeerTailEntries entries per rank, laid out [table][rank][entry] = the global table in blocked order.
constexpr int kPeerTailTables = 8;
constexpr size_t kPeerTailEntries = 1024;
constexpr size_t kPeerTailTableStride = (size_t)kPeerMaxWorld * kPeerTailEntries;             // entries
constexpr size_t kPeerTailBufBytes = (size_t)kPeerTailTables * kPeerTailTableStride * 32;   // 2 MiB
constexpr size_t kPeerBoxBytes = kPeerHeaderBytes + 2 * kPeerTailBufBytes;
static_assert(kPeerDataBytes + 2ull * kPeerMaxWorld * sizeof(unsigned) <= kPeerHeaderBytes, "mailbox header");
constexpr unsigned long long kPeerTimeoutNsDefault = 4000000000ull;  // ZOLT_TUNE_PEER_TIMEOUT_MS overrides it per context

struct PeerDev {
  unsigned char* box[kPeerMaxWorld];  // every rank's mailbox in this process's address space (box[rank] is local)
  int rank, world;                    // world <= 1: no exchange
  unsigned seq;                       // this exchange's number (> 0)
  int* err;                           // mapped host flag, set when a peer never showed up
  unsigned long long timeout_ns;      // how long a kernel waits for its peers, measured from ITS OWN start: ranks must be
                                      // roughly aligned (zolt_peer_barrier after set-up) or the knob raised
};

#ifdef __CUDACC__
__device__ __forceinline__ Fr* peer_row(unsigned char* base, unsigned seq, int sender) {
  return reinterpret_cast<Fr*>(base) + ((size_t)(seq & 1u) * kPeerMaxWorld + sender) * kPeerSlotFr;
}
__device__ __forceinline__ unsigned* peer_flag(unsigned char* base, unsigned seq, int sender) {
  return reinterpret_cast<unsigned*>(base + kPeerDataBytes) + (seq & 1u) * kPeerMaxWorld + sender;
}

// One thread: store n_fr elements into every rank's mailbox row of this rank, then raise the flags.
static __device__ __noinline__ void peer_publish(const PeerDev& p, const Fr* vals, int n_fr) {
#pragma unroll 1
  for (int g = 0; g < p.world; g++) {
    uint4* dst = reinterpret_cast<uint4*>(peer_row(p.box[g], p.seq, p.rank));
#pragma unroll 1
    for (int i = 0; i < n_fr; i++) {
      dst[2 * i] = make_uint4(vals[i].v[0], vals[i].v[1], vals[i].v[2], vals[i].v[3]);
      dst[2 * i + 1] = make_uint4(vals[i].v[4], vals[i].v[5], vals[i].v[6], vals[i].v[7]);
    }
  }
  // ONE system-scope fence orders all the row stores above before all the flag stores below; the flags themselves
  // are relaxed stores (a release per flag would pay the NVLink round trip once per peer)
  __threadfence_system();
#pragma unroll 1
  for (int g = 0; g < p.world; g++) {
    unsigned* f = peer_flag(p.box[g], p.seq, p.rank);
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(f), "r"(p.seq) : "memory");
  }
}

// One thread: wait until every rank's row of this exchange has landed in the local mailbox.
static __device__ __noinline__ bool peer_wait(const PeerDev& p) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
#pragma unroll 1
  for (int g = 0; g < p.world; g++) {
    const unsigned* f = peer_flag(p.box[p.rank], p.seq, g);
    unsigned spins = 0;
    while (true) {
      unsigned v;
      asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if (v == p.seq) break;
      if ((++spins & 1023u) == 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > p.timeout_ns) {
          *p.err = 1 + g;
          __threadfence_system();
          return false;
        }
      }
    }
  }
  __threadfence_system();  // acquire side: the rows are read (from L2) only after every flag has been seen
  return true;
}

// element i of sender g's row in the local mailbox (L2 load: the row was written over NVLink)
__device__ __forceinline__ Fr peer_load(const PeerDev& p, int g, int i) {
  const uint4* q = reinterpret_cast<const uint4*>(peer_row(p.box[p.rank], p.seq, g) + i);
  const uint4 a = __ldcg(q), b = __ldcg(q + 1);
  Fr x;
  x.v[0] = a.x; x.v[1] = a.y; x.v[2] = a.z; x.v[3] = a.w;
  x.v[4] = b.x; x.v[5] = b.y; x.v[6] = b.z; x.v[7] = b.w;
  return x;
}
#endif

}  // namespace zolt

// Host-side handle (zolt_peer_box in include/zolt_gpu.h)
struct zolt_peer_box {
  zolt_ctx* ctx = nullptr;
  unsigned char* d_local = nullptr;
  unsigned char* mapped[zolt::kPeerMaxWorld] = {};
  bool opened[zolt::kPeerMaxWorld] = {};
  int rank = 0, world = 1;
  unsigned seq = 0;
  unsigned gathers = 0;  // table gathers so far (selects the tail buffer)
  int* h_err = nullptr;  // mapped pinned
  int* d_err = nullptr;
  bool connected = false;
};

namespace zolt {
// PeerDev for the next exchange (advances the exchange counter); world 1 when box is null.
inline PeerDev peer_next(zolt_peer_box* box) {
  PeerDev p{};
  p.world = 1;
  if (!box || !box->connected) return p;
  for (int g = 0; g < box->world; g++) p.box[g] = box->mapped[g];
  p.rank = box->rank;
  p.world = box->world;
  p.seq = ++box->seq;
  if (p.seq == 0) p.seq = box->seq = 2;  // wrap: keep the slot parity alternating and the number non-zero
  p.err = box->d_err;
  p.timeout_ns = box->ctx && box->ctx->peer_timeout_ms > 0 ? (unsigned long long)box->ctx->peer_timeout_ms * 1000000ull : kPeerTimeoutNsDefault;
  return p;
}
// peer.cu
int peer_allgather_ptrs(zolt_peer_box* box, const Fr* const* d_src, int n_fr, Fr* d_out, int transpose);
// All-gather k local tables of n_local (<= kPeerTailEntries) entries each: afterwards d_tables_out[t] points at the
// global table t (world * n_local entries, rank-major = blocked order) inside THIS rank's mailbox.  Asynchronous.
int peer_gather_tables(zolt_peer_box* box, const Fr* const* d_src, int k, size_t n_local, const Fr** d_tables_out);
int peer_check(zolt_peer_box* box);
}  // namespace zolt
