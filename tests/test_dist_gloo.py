"""CPU tier, world_size 2 over gloo: the host-side logic of the multi-GPU path (zolt_b200/dist.py) --
blocked hypercube sharding, per-round all-gather + fold of partial evaluations, transcript coupling,
the tail gather -- with a big-integer backend standing in for the GPU kernels.  The sharded run must
reproduce the single-process protocol bit for bit."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import bn254 as B  # noqa: E402
from oracle import cpu_ref as R  # noqa: E402
from oracle import sumcheck as OS  # noqa: E402
from oracle import transcript as OT  # noqa: E402
from tests.util import P, fr_arr, fr_ints, rand_fr  # noqa: E402
from zolt_b200 import dist as ZD  # noqa: E402


class IntBackend:
    """Backend protocol of run_sharded_sumcheck over Python integers + gloo."""

    def __init__(self, tables, world, kind):
        self.tables, self.world, self.kind = [list(t) for t in tables], world, kind

    def local_evals(self):
        e = OS.ROUND_FNS[self.kind](self.tables, OS.ADJACENT)
        if self.world > 1:
            mine = torch.from_numpy(fr_arr(e).view(np.int64).copy())
            allv = [torch.zeros_like(mine) for _ in range(self.world)]
            dist.all_gather(allv, mine)
            parts = [fr_ints(t.numpy().view(np.uint64)) for t in allv]
            e = [sum(p[j] for p in parts) % P for j in range(len(e))]
        return fr_arr(e)

    def bind(self, r):
        self.tables = OS.bind_tables(self.tables, fr_ints(r)[0], OS.ADJACENT)

    def gather_tail(self):
        mine = torch.from_numpy(fr_arr([t[0] for t in self.tables]).view(np.int64).copy())
        allv = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(allv, mine)
        cols = [fr_ints(t.numpy().view(np.uint64)) for t in allv]  # per rank: k values
        tables = [[cols[g][t] for g in range(self.world)] for t in range(len(self.tables))]
        return IntBackend(tables, 1, self.kind)

    def final_values(self):
        return fr_arr([t[0] for t in self.tables])


class PyTranscript:
    def __init__(self, domain):
        self.t = OT.Transcript(domain)

    def append_scalar(self, label, limbs):
        self.t.append_scalar(label, fr_ints(limbs)[0])

    def challenge_scalar(self, label):
        return fr_arr([self.t.challenge_scalar(label)])[0]


def _worker(rank, world, port, log_n, kind, k, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = [rand_fr(100 + t, 1 << log_n) for t in range(k)]
    lo, hi = ZD.shard_range(1 << log_n, rank, world)
    backend = IntBackend([t[lo:hi] for t in full], world, kind)
    n_absorb = {"sum": 2, "prod2": 3}.get(kind, 4)
    evals, chal, final = ZD.run_sharded_sumcheck(backend, PyTranscript(b"bench"), log_n - ZD.log2_exact(world), world, n_absorb)
    # MSM control flow: partial per point range -> all-gather -> combine
    logs, scalars = rand_fr(7, 64), rand_fr(8, 64)
    mlo, mhi = ZD.shard_range(64, rank, world)

    def partial():
        return sum(a * b for a, b in zip(logs[mlo:mhi], scalars[mlo:mhi])) % P

    def allgather(p):
        mine = torch.from_numpy(fr_arr([p]).view(np.int64).copy())
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        return [fr_ints(t.numpy().view(np.uint64))[0] for t in allv]

    msm = ZD.run_sharded_msm(partial, allgather, lambda parts: B.g1_mul(B.G1_GEN, sum(parts)))
    if rank == 0:
        q.put((evals, chal, final, msm))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind,k", [("eq_ab_minus_c", 4), ("prod2", 2), ("sum", 1)])
def test_sharded_protocol_world2(kind, k):
    log_n, world = 6, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + hash(kind)) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, log_n, kind, k, q)) for r in range(world)]
    for p in procs:
        p.start()
    evals, chal, final, msm = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = [rand_fr(100 + t, 1 << log_n) for t in range(k)]
    r_evals, r_chal, r_final = R.sc_prove(kind, 0, [fr_arr(t) for t in full])
    ne = {"sum": 2, "prod2": 3}.get(kind, 4)
    assert (chal == r_chal).all()
    assert (evals[:, :ne] == r_evals[:, :ne]).all()
    assert all((final[t] == r_final[t][0]).all() for t in range(k))
    assert msm == B.msm_closed_form(rand_fr(7, 64), rand_fr(8, 64))


def test_shard_helpers():
    assert ZD.shard_range(16, 1, 4) == (4, 8)
    assert ZD.log2_exact(8) == 3
    with pytest.raises(ValueError):
        ZD.shard_range(10, 0, 4)
    with pytest.raises(ValueError):
        ZD.log2_exact(6)
