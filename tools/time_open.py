"""Time HyperKZG.open on device-resident tables (not a benchmark line; numbers quoted in DESIGN.md).
    python tools/time_open.py [log_n=24]"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import zolt_b200 as Z  # noqa: E402


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    n = 1 << log_n
    ctx = Z.Context(0)
    rng = np.random.default_rng(1)
    ev = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    ev[:, 3] &= np.uint64((1 << 60) - 1)
    pt = rng.integers(0, 1 << 64, size=(log_n, 4), dtype=np.uint64)
    pt[:, 3] &= np.uint64((1 << 60) - 1)
    srs = Z.HyperKZG.setup(ctx, n).precompute(0)
    d = Z.DeviceBuffer.from_host(ctx, ev)
    for label in ("warm-up", "timed", "timed"):
        ctx.sync()
        t0 = time.perf_counter()
        q, inf, fin = Z.HyperKZG.open_dev(srs, d.ptr, n, pt)
        dt = (time.perf_counter() - t0) * 1e3
        print(f"open 2^{log_n} ({label}): {dt:.2f} ms, {len(q)} quotient commitments, launches so far {ctx.launch_count}")
    t0 = time.perf_counter()
    Z.MSM.compute_dev(srs, d.ptr, n)
    print(f"commit 2^{log_n}: {(time.perf_counter() - t0) * 1e3:.2f} ms")


if __name__ == "__main__":
    main()
