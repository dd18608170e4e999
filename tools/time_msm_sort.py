"""A/B of the MSM bucket sort (two-pass bins vs atomic scatter, tuning key 9); diagnostics, not a benchmark of record.

    python tools/time_msm_sort.py [log_n=24]
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import zolt_b200 as Z  # noqa: E402

TAGS = {1: "digits/partition", 2: "scan/count", 3: "scatter/place", 4: "accum0", 5: "edges", 6: "reduce", 7: "affine"}


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    n = 1 << log_n
    ctx = Z.Context(0)
    rng = np.random.default_rng(5)
    k = rng.integers(1, 1 << 62, size=(n, 4), dtype=np.uint64)
    k[:, 3] &= np.uint64((1 << 59) - 1)
    bases = Z.Bases.fixed_base_mul(ctx, k)
    bases.precompute(0)
    s = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    s[:, 3] &= np.uint64((1 << 60) - 1)
    small = s.copy()
    small[:, 1:] = 0
    small[:, 0] &= np.uint64(255)
    small[::3] = 0
    for name, sc in (("uniform", s), ("bytes", ctx.field_op(Z.FIELD_FR, Z.OP_TO_MONT, small) if n <= (1 << 22) else None)):
        if sc is None:
            continue
        d_s = Z.DeviceBuffer.from_host(ctx, sc)
        res = {}
        for mode in (1, 0):
            ctx.set_tuning(9, mode)
            Z.MSM.compute_dev(bases, d_s.ptr, n)
            best = None
            for _ in range(3):
                ctx.prof_enable(True)
                ctx.timer_start()
                out = Z.MSM.compute_dev(bases, d_s.ptr, n)
                wall = ctx.timer_stop_ms()
                log = ctx.prof_read()
                ctx.prof_enable(False)
                if best is None or wall < best[0]:
                    best = (wall, log)
            res[mode] = out
            ph = {}
            for tag, ms in best[1]:
                ph[TAGS.get(tag, tag)] = ph.get(TAGS.get(tag, tag), 0) + ms
            print(name, "two-pass" if mode == 0 else "atomic scatter", f"{best[0]:.3f} ms", {k2: round(v, 3) for k2, v in ph.items()}, flush=True)
        same = (res[0][0] == res[1][0]).all() and res[0][1] == res[1][1]
        print(name, "same point:", bool(same))
        d_s.free()


if __name__ == "__main__":
    main()
