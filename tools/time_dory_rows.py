"""Time the Dory row-commitment shape (dory.zig:646-670): k MSMs of n points over shared bases (zolt_msm_batch).
    python tools/time_dory_rows.py [log_k=12] [log_n=12] [precompute window bits, 0 = plain bases]"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import zolt_b200 as Z  # noqa: E402
from oracle import bn254 as B  # noqa: E402  (checker only)


def main():
    log_k = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    log_n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    pre_c = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    k, n = 1 << log_k, 1 << log_n
    ctx = Z.Context(0)
    rng = np.random.default_rng(1)
    logs = rng.integers(1, 1 << 62, size=n, dtype=np.uint64)
    lg = np.zeros((n, 4), dtype=np.uint64)
    lg[:, 0] = logs
    bases = Z.Bases.fixed_base_mul(ctx, ctx.field_op(Z.FIELD_FR, Z.OP_TO_MONT, lg))
    if pre_c:
        bases.precompute(pre_c)
    sc = rng.integers(0, 1 << 64, size=(k * n, 4), dtype=np.uint64)
    sc[:, 3] &= np.uint64((1 << 60) - 1)
    for label in ("warm-up", "timed", "timed"):
        t0 = time.perf_counter()
        xy, inf = Z.MSM.compute_batch(bases, sc, k, n)
        print(f"batch {k} x {n} pre_c={pre_c} ({label}): {(time.perf_counter() - t0) * 1e3:.1f} ms  (includes H2D of {sc.nbytes >> 20} MiB)")
    tags = {1: "digits", 2: "scan", 3: "scatter", 4: "accum0", 5: "edges", 6: "reduce", 7: "affine"}
    ctx.prof_enable(True)
    Z.MSM.compute_batch(bases, sc, k, n)
    print({tags.get(t, t): round(ms, 3) for t, ms in ctx.prof_read()})
    ctx.prof_enable(False)
    # spot-check row 0 and the last row against single MSMs
    for row in (0, k - 1):
        one = Z.MSM.compute(bases, sc[row * n:(row + 1) * n])
        assert (one[0] == xy[row]).all() and one[1] == inf[row]
    print("rows 0 and k-1 equal the single-MSM results")


if __name__ == "__main__":
    main()
