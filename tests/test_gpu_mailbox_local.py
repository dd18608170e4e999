"""The NVLink mailbox protocol (csrc/peer.cuh) on ONE GPU: 2, 4 and 8 ranks = that many contexts with their own streams in one process,
mailboxes addressed directly (zolt_peer_box_connect_local), each rank driven by its own host thread.  Exercises exactly the
kernels and the C++ round loop the multi-GPU runs use (publish / flag / wait inside the round kernels, the MSM partial
sum, the table gather of the tail), so the exchange code is covered on a single-GPU box as well; tests/test_gpu_multi.py
is the same check over real peers (2 GPUs, CUDA IPC)."""
import threading

import numpy as np
import pytest

import zolt_b200 as Z
from oracle import bn254 as B
from oracle import cpu_ref as R
from tests.util import fr_arr, rand_fr
from zolt_b200 import dist as ZD
from zolt_b200 import host as H

pytestmark = pytest.mark.gpu


def _run_ranks(fn, WORLD):
    """fn(rank, barrier) on one thread per rank; re-raises the first failure.

    Two ranks on ONE device share its allocator: cudaMalloc / cudaFree synchronise the whole device, so a rank that allocates
    while its peer's kernel is already spinning on the mailbox would stall until the kernel's time-out.  (Ranks on separate
    GPUs, the real deployment, do not interact that way.)  Each rank therefore warms its buffer pool up first and the
    ranks meet at `barrier` before the exchange and again before anything is freed."""
    out, err = [None] * WORLD, []
    barrier = threading.Barrier(WORLD, timeout=60)

    def work(rank):
        try:
            out[rank] = fn(rank, barrier)
        except BaseException as e:  # noqa: BLE001
            err.append(e)
            barrier.abort()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(WORLD)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in ts), "a rank is stuck"
    if err:
        raise err[0]
    return out


@pytest.fixture(scope="module", params=[2, 4, 8], ids=lambda w: f"world{w}")
def ranks(built_lib, request):
    WORLD = request.param
    ctxs = [Z.Context(0) for _ in range(WORLD)]  # private non-blocking streams: the ranks' kernels overlap on the GPU
    boxes = Z.PeerBox.local_group(ctxs)
    yield ctxs, boxes
    for b in boxes:
        b.free()
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("kind,comb,k,log_n", [("eq_ab_minus_c", Z.SC_EQ_AB_MINUS_C, 4, 13), ("prod2", Z.SC_PROD2, 2, 16),
                                               ("eq_ab_minus_c", Z.SC_EQ_AB_MINUS_C, 4, 5)])
def test_sharded_sumcheck_ranks_one_gpu(ranks, kind, comb, k, log_n):
    ctxs, boxes = ranks
    WORLD = len(ctxs)
    local_log = log_n - (WORLD.bit_length() - 1)
    full = [fr_arr(rand_fr(700 + 10 * log_n + t, 1 << log_n)) for t in range(k)]
    r_evals, r_chal, r_final = R.sc_prove(kind, 0, full)
    ne = 3 if comb == Z.SC_PROD2 else 4

    def rank_fn(rank, barrier):
        lo, hi = ZD.shard_range(1 << log_n, rank, WORLD)
        ctx = ctxs[rank]
        bufs = [Z.DeviceBuffer.from_host(ctx, t[lo:hi]) for t in full]
        H.sumcheck_prove_dev(ctx, comb, Z.PAIR_ADJACENT, [b.ptr for b in bufs], local_log)  # warm-up: fills the buffer pool
        ctx.sync()
        barrier.wait()
        res = H.sumcheck_prove_sharded(ctx, boxes[rank], comb, [b.ptr for b in bufs], local_log)
        ctx.sync()
        barrier.wait()
        for b in bufs:
            b.free()
        return res

    for evals, chal, final in _run_ranks(rank_fn, WORLD):
        assert (chal == r_chal).all() and (evals[:, :ne] == r_evals[:, :ne]).all()
        assert all((final[t] == r_final[t][0]).all() for t in range(k))


def test_sharded_msm_and_primitives_ranks_one_gpu(ranks):
    ctxs, boxes = ranks
    WORLD = len(ctxs)
    n = 1 << 12
    logs, scalars = rand_fr(811, n), rand_fr(812, n)
    want = B.msm_closed_form(logs, scalars)

    def rank_fn(rank, barrier):
        lo, hi = ZD.shard_range(n, rank, WORLD)
        ctx = ctxs[rank]
        bases = Z.Bases.fixed_base_mul(ctx, fr_arr(logs[lo:hi]))
        d_s = Z.DeviceBuffer.from_host(ctx, fr_arr(scalars[lo:hi]))
        src = Z.DeviceBuffer.from_host(ctx, np.arange(rank * 16, rank * 16 + 16, dtype=np.uint64).reshape(4, 4))
        out = Z.DeviceBuffer(ctx, 16 * WORLD * 8)
        sc_host = fr_arr(scalars[lo:hi])
        Z.MSM.compute_dev(bases, d_s.ptr, hi - lo)  # warm-up: MSM workspace and pool buffers exist before anybody spins
        Z.MSM.compute(bases, sc_host)               # ... and the staging buffer of the host-scalar path
        ctx.sync()
        barrier.wait()
        xy, inf = Z.MSM.compute_sharded_dev(bases, d_s.ptr, hi - lo, boxes[rank])
        xy_h, inf_h = Z.MSM.compute_sharded(bases, sc_host, boxes[rank])  # host scalars, same exchange
        assert (xy_h == xy).all() and inf_h == inf
        # all-gather primitive (both layouts) and the stream-ordered barrier
        boxes[rank].allgather(src.ptr, 4, out.ptr, False)
        flat = out.download((16 * WORLD,)).tolist()
        boxes[rank].allgather(src.ptr, 4, out.ptr, True)
        tr = out.download((16 * WORLD,)).tolist()
        boxes[rank].barrier()
        ctx.sync()
        barrier.wait()
        for b in (d_s, src, out):
            b.free()
        bases.free()
        return xy, inf, flat, tr

    for xy, inf, flat, tr in _run_ranks(rank_fn, WORLD):
        assert B.arrays_to_points(xy.reshape(1, 8), [inf])[0] == want
        assert flat == list(range(16 * WORLD))
        assert tr == [g * 16 + i * 4 + j for i in range(4) for g in range(WORLD) for j in range(4)]
