"""Determinism + oracle check of the streaming sumcheck kernels at a large size (diagnostic)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import zolt_b200 as Z
from zolt_b200 import host as H
from oracle import cpu_ref as R

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
masks = [int(x, 0) for x in (sys.argv[2] if len(sys.argv) > 2 else "0,0,0x800,0xf,0x80f").split(",")]
ctx = Z.Context(0)
rng = np.random.default_rng(5)
n = 1 << log_n
tabs = []
for _ in range(4):
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 60) - 1)
    tabs.append(a)
bufs = [Z.DeviceBuffer.from_host(ctx, a) for a in tabs]
r_evals, r_chal, r_final = R.sc_prove("eq_ab_minus_c", 0, tabs)
for m in masks:
    ctx.set_tuning(2, m)
    ev, ch, fin = H.sumcheck_prove_dev(ctx, Z.SC_EQ_AB_MINUS_C, Z.PAIR_ADJACENT, [b.ptr for b in bufs], log_n)
    ok_e = (ev == r_evals).all(axis=(1, 2))
    first_bad = int(np.argmin(ok_e)) if not ok_e.all() else -1
    ok_f = all((fin[t] == r_final[t][0]).all() for t in range(4))
    print(f"mask {m:#x}: evals ok={bool(ok_e.all())} first bad round={first_bad} final ok={ok_f}", flush=True)
ctx.close()
