"""Timing of the three Stage 3 provers on the device (not a benchmark of record).

    python tools/time_stage3.py [log_T ...]
16 trace columns of T cycles; prints wall per proof (table construction, input claims, log T rounds, final claims) for every
size given, so the size-independent part (per-round latency of seven instances) separates from the streaming part.
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import zolt_b200 as Z  # noqa: E402
from zolt_b200 import host as H  # noqa: E402


def rand_limbs(rng, n):
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [12, 16, 18, 20, 22]
    ctx = Z.Context(0)
    rng = np.random.default_rng(3)
    out = {}
    for log_T in sizes:
        T = 1 << log_T
        base = rand_limbs(rng, T)
        bufs = []
        for k in range(13):
            col = base.copy()
            col[:, 0] += np.uint64(k)
            bufs.append(Z.DeviceBuffer.from_host(ctx, col))
        # STAGE3_COLUMNS order; instr.unexpanded_pc / reg.rs1_value / reg.rs2_value share the columns of the same name
        order = [0, 1, 2, 3, 4, 5, 6, 7, 0, 8, 9, 10, 11, 12, 6, 9]
        cols = [bufs[i].ptr for i in order]
        args = (rand_limbs(rng, log_T), rand_limbs(rng, log_T), rand_limbs(rng, 7), rand_limbs(rng, 3), rand_limbs(rng, log_T))
        best = None
        for rep in range(2):
            ctx.prof_enable(True)
            ctx.timer_start()
            H.stage3_run(ctx, log_T, cols, *args)
            wall = ctx.timer_stop_ms()
            ks = [ms for tag, ms in ctx.prof_read() if tag in (16, 17, 18)]
            ctx.prof_enable(False)
            if rep and (best is None or wall < best[0]):
                best = (wall, ks)
        out[f"T2^{log_T}"] = {"wall_ms": round(best[0], 3), "round_launches": len(best[1]), "round_kernel_ms_sum": round(sum(best[1]), 3),
                             "largest_launches_ms": [round(x, 3) for x in sorted(best[1], reverse=True)[:12]]}
        for b in bufs:
            b.free()
    print(json.dumps(out))
    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "time_stage3.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
