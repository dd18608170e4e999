"""Multi-GPU sharding of the hot path, one process per GPU.

  * MSM: shard by point range (the reference's ParallelMSM split, src/msm/mod.zig:608-653).  Each rank
    computes the partial sum of its range; ONE all-gather of a 128-byte partial per rank; every rank
    adds the partials on its device and normalises.
  * Sumcheck, adjacent-pair (LowToHigh) order: BLOCKED sharding -- rank g owns the contiguous index
    range [g*N/G, (g+1)*N/G), so pairs (2i, 2i+1) are always local and after a bind the shard is the
    contiguous range of the halved table.  Per round: local partial evaluations -> all-gather
    (G x 128 B) -> folded mod r -> host transcript (every rank derives the same challenge) -> local
    fused bind+eval.  When the local shard is down to one entry per table the G remaining entries are
    gathered and the last log2(G) rounds run redundantly on every rank.

The round driver is written against a small backend protocol so the host-side logic (round
bookkeeping, transcript coupling, the tail gather) is testable on CPU with gloo and a big-integer
backend (tests/test_dist_gloo.py); GpuSumcheckBackend is the product backend.
"""
from __future__ import annotations

import numpy as np

from . import _lib


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of rank `rank` (n divisible by world)."""
    if n % world:
        raise ValueError("table length must be divisible by the world size")
    per = n // world
    return rank * per, (rank + 1) * per


def log2_exact(v: int) -> int:
    if v <= 0 or v & (v - 1):
        raise ValueError("power of two expected")
    return v.bit_length() - 1


def run_sharded_sumcheck(backend, transcript, local_log: int, world: int, n_absorb: int, label: bytes = b"round"):
    """Drive all rounds.  backend protocol:
         local_evals()            -> (ne, 4) uint64 evaluations summed over ALL ranks (does the exchange)
         bind(r)                  -> bind the local shard to challenge r ((4,) uint64)
         gather_tail()            -> backend for the last log2(world) rounds over the gathered tables
       transcript: append_scalar(label, limbs), challenge_scalar(label) -> (4,) uint64
    Returns (evals (rounds, ne, 4), challenges (rounds, 4), final table values)."""
    tail_log = log2_exact(world)
    evals_all, chal_all = [], []

    def one_round(b):
        e = b.local_evals()
        for j in range(n_absorb):
            transcript.append_scalar(b"round_poly_%d" % j, e[j])
        r = transcript.challenge_scalar(label)
        b.bind(r)
        evals_all.append(np.array(e, dtype=np.uint64))
        chal_all.append(np.array(r, dtype=np.uint64))

    for _ in range(local_log):
        one_round(backend)
    tail = backend.gather_tail() if world > 1 else backend
    for _ in range(tail_log):
        one_round(tail)
    return np.stack(evals_all) if evals_all else np.zeros((0, 4, 4), np.uint64), np.stack(chal_all) if chal_all else np.zeros((0, 4), np.uint64), tail.final_values()


class GpuSumcheckBackend:
    """Product backend: SumcheckInstance on this rank's GPU + NCCL all-gather of 128-byte partials."""

    def __init__(self, ctx, stream, inst, world: int, ev_part, ev_gath, combiner: int, k: int):
        self.ctx, self.stream, self.inst, self.world = ctx, stream, inst, world
        self.ev_part, self.ev_gath = ev_part, ev_gath
        self.combiner, self.k = combiner, k
        self.ne = inst.num_evals
        if world > 1:
            inst.set_sharded(True)

    def local_evals(self):
        import torch
        import torch.distributed as dist

        if self.world == 1:
            return self.inst.compute_round_poly()
        d_ptr = self.inst.partials_dev()  # this rank's 4 raw partial sums, left on the device
        with torch.cuda.stream(self.stream):
            # stage into the torch tensor NCCL gathers from (device-to-device, 128 bytes)
            _lib.check(self.ctx.lib.zolt_dev_copy(self.ctx.handle, self.ev_part.data_ptr(), d_ptr, 128), "zolt_dev_copy")
            dist.all_gather_into_tensor(self.ev_gath, self.ev_part)
            # fold the rows mod r and complete the round polynomial on the device (every rank alike)
            return self.inst.finish_partials(self.ev_gath.data_ptr(), self.world)

    def bind(self, r):
        self.inst.bind_challenge(r)

    def gather_tail(self):
        import torch
        import torch.distributed as dist

        from .api import SumcheckInstance

        vals = np.stack([self.inst.table(t)[0] for t in range(self.k)])  # (k, 4): one entry per table left
        with torch.cuda.stream(self.stream):
            mine = torch.from_numpy(vals.view(np.int64).copy()).cuda()
            allv = torch.zeros((self.world,) + tuple(mine.shape), dtype=mine.dtype, device="cuda")
            dist.all_gather_into_tensor(allv, mine)
            host = allv.cpu().numpy().view(np.uint64)  # (world, k, 4); index = rank = remaining variables
        tables = [np.ascontiguousarray(host[:, t, :]) for t in range(self.k)]
        from ._lib import PAIR_ADJACENT

        tail = SumcheckInstance(self.ctx, self.combiner, PAIR_ADJACENT, tables)
        return GpuSumcheckBackend(self.ctx, self.stream, tail, 1, None, None, self.combiner, self.k)

    def final_values(self):
        return np.stack([self.inst.table(t)[0] for t in range(self.k)])


def sharded_sumcheck_gpu(ctx, stream, combiner: int, d_tab_ptrs, local_log: int, world: int, ev_part, ev_gath,
                         domain: bytes = b"bench"):
    from . import host as H
    from ._lib import PAIR_ADJACENT, SC_PROD2, SC_SUM
    from .api import SumcheckInstance

    inst = SumcheckInstance.from_device(ctx, combiner, PAIR_ADJACENT, d_tab_ptrs, local_log)
    backend = GpuSumcheckBackend(ctx, stream, inst, world, ev_part, ev_gath, combiner, len(d_tab_ptrs))
    n_absorb = 2 if combiner == SC_SUM else (3 if combiner == SC_PROD2 else 4)
    out = run_sharded_sumcheck(backend, H.Transcript(domain), local_log, world, n_absorb)
    inst.free()
    return out


def connect_peer_box(ctx, rank: int, world: int):
    """Create this rank's NVLink mailbox and map every peer's (one torch.distributed all_gather of the 64-byte
    CUDA IPC handles at start-up; nothing on the data path uses the process group afterwards)."""
    import torch
    import torch.distributed as dist

    from .api import PeerBox

    def exchange(blob: bytes):
        mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda()
        allb = torch.zeros(64 * world, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(allb, mine)
        host = bytes(allb.cpu().numpy().tobytes())
        return [host[64 * g:64 * (g + 1)] for g in range(world)]

    box, err = None, None
    try:
        box = PeerBox(ctx, rank, world, exchange)
    except Exception as e:  # noqa: BLE001 - reported below, on every rank alike
        err = e
    # every rank has mapped every mailbox before anybody publishes -- or nobody uses the mailboxes at all
    ok = torch.tensor([0 if err is not None else 1], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) == 0:
        if box is not None:
            box.free()
        raise RuntimeError(f"peer mailboxes unavailable on at least one rank ({err!r} on rank {rank})")
    return box


def run_sharded_msm(partial_fn, allgather_fn, combine_fn):
    """MSM over point-range shards: partial_fn() -> this rank's partial; allgather_fn(p) -> all partials;
    combine_fn(parts) -> affine result.  (Kept as three callables so the CPU gloo test can exercise the
    same control flow with a big-integer backend.)"""
    return combine_fn(allgather_fn(partial_fn()))
