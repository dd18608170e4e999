"""GPU parity, 2 / 4 / 8 ranks, one process per GPU (each case skipped unless that many GPUs are visible): the sharded MSM and sharded
sumcheck of zolt_b200/dist.py must reproduce the single-process results bit for bit."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["ZOLT_ROOT"])
import zolt_b200 as Z
from zolt_b200 import dist as ZD
from oracle import bn254 as B, cpu_ref as R
from tests.util import P, fr_arr, fr_ints, rand_fr

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
stream = torch.cuda.Stream()
ctx = Z.Context(rank, stream.cuda_stream)

# --- MSM sharded by point range -------------------------------------------------------------------
n = 1 << 13
logs, scalars = rand_fr(11, n), rand_fr(12, n)
lo, hi = ZD.shard_range(n, rank, world)
bases = Z.Bases.fixed_base_mul(ctx, fr_arr(logs[lo:hi]))
d_s = Z.DeviceBuffer.from_host(ctx, fr_arr(scalars[lo:hi]))
part = torch.zeros(128, dtype=torch.uint8, device="cuda")
gathered = torch.zeros(128 * world, dtype=torch.uint8, device="cuda")
with torch.cuda.stream(stream):
    Z.MSM.partial_dev(bases, d_s.ptr, hi - lo, part.data_ptr())
    dist.all_gather_into_tensor(gathered, part)
    xy, inf = Z.MSM.sum_partials(ctx, gathered.data_ptr(), world)
got = B.arrays_to_points(xy.reshape(1, 8), [inf])[0]
assert got == B.msm_closed_form(logs, scalars), "sharded MSM differs"

# --- sumcheck sharded by contiguous hypercube blocks ---------------------------------------------------
log_n, k = 12, 4
full = [fr_arr(rand_fr(100 + t, 1 << log_n)) for t in range(k)]
lo, hi = ZD.shard_range(1 << log_n, rank, world)
d_tabs = [Z.DeviceBuffer.from_host(ctx, t[lo:hi]) for t in full]
ev_part = torch.zeros(128, dtype=torch.uint8, device="cuda")
ev_gath = torch.zeros(128 * world, dtype=torch.uint8, device="cuda")
evals, chal, final = ZD.sharded_sumcheck_gpu(ctx, stream, Z.SC_EQ_AB_MINUS_C, [b.ptr for b in d_tabs],
                                             log_n - ZD.log2_exact(world), world, ev_part, ev_gath)
r_evals, r_chal, r_final = R.sc_prove("eq_ab_minus_c", 0, full)
assert (chal == r_chal).all(), "challenges differ"
assert (evals == r_evals).all(), "round evaluations differ"
assert all((final[t] == r_final[t][0]).all() for t in range(k)), "final values differ"

# --- the same two paths with the exchange fused into the kernels (NVLink peer mailboxes, no NCCL on the data path) ---
from zolt_b200 import host as H
box = ZD.connect_peer_box(ctx, rank, world)
xy2, inf2 = Z.MSM.compute_sharded_dev(bases, d_s.ptr, (1 << 13) // world, box)
assert (xy2 == xy).all() and inf2 == inf, "mailbox-sharded MSM differs"
xy3, inf3 = Z.MSM.compute_sharded(bases, fr_arr(scalars[(1 << 13) // world * rank:(1 << 13) // world * (rank + 1)]), box)
assert (xy3 == xy).all() and inf3 == inf, "mailbox-sharded MSM from host scalars differs"
e2, c2, f2 = H.sumcheck_prove_sharded(ctx, box, Z.SC_EQ_AB_MINUS_C, [b.ptr for b in d_tabs], log_n - ZD.log2_exact(world))
assert (c2 == r_chal).all() and (e2 == r_evals).all(), "mailbox-sharded sumcheck: round polynomials differ"
assert all((f2[t] == r_final[t][0]).all() for t in range(k)), "mailbox-sharded sumcheck: final values differ"
# larger block: streaming round kernels first, then the small-round kernel, then the gathered tail; degree 2 and 3
for comb, name, kk, lg in ((Z.SC_EQ_AB_MINUS_C, "eq_ab_minus_c", 4, 17), (Z.SC_PROD2, "prod2", 2, 16)):
    full = [fr_arr(rand_fr(500 + 10 * lg + t, 1 << lg)) for t in range(kk)]
    lo, hi = ZD.shard_range(1 << lg, rank, world)
    d_big = [Z.DeviceBuffer.from_host(ctx, t[lo:hi]) for t in full]
    e3, c3, f3 = H.sumcheck_prove_sharded(ctx, box, comb, [b.ptr for b in d_big], lg - ZD.log2_exact(world))
    r_e, r_c, r_f = R.sc_prove(name, 0, full)
    ne = 3 if comb == Z.SC_PROD2 else 4
    assert (c3 == r_c).all() and (e3[:, :ne] == r_e[:, :ne]).all(), f"mailbox-sharded {name} 2^{lg}: round polynomials differ"
    assert all((f3[t] == r_f[t][0]).all() for t in range(kk)), f"mailbox-sharded {name} 2^{lg}: final values differ"
    for b in d_big:
        b.free()
# all-gather primitive, both layouts
src = torch.arange(rank * 16, rank * 16 + 16, dtype=torch.int64, device="cuda")  # 4 field-sized elements
out = torch.zeros(16 * world, dtype=torch.int64, device="cuda")
box.allgather(src.data_ptr(), 4, out.data_ptr(), False)
assert out.cpu().tolist() == list(range(16 * world))
box.allgather(src.data_ptr(), 4, out.data_ptr(), True)
exp = [g * 16 + i * 4 + j for i in range(4) for g in range(world) for j in range(4)]
assert out.cpu().tolist() == exp
dist.barrier()
box.free()
if rank == 0:
    print("MULTI_GPU_OK")
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_rank_parity(tmp_path, world):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (run under `gpurun --gpus {world}`)")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, ZOLT_ROOT=str(ROOT))
    res = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
         "--master-port", str(29533 + world), str(script)],
        capture_output=True, text=True, env=env, timeout=600,
    )
    assert res.returncode == 0 and "MULTI_GPU_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
