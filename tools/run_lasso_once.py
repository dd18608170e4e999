"""Run one LassoProver proof on the device (for `ncu -k regex:lasso_addr`); not a benchmark.

    python tools/run_lasso_once.py [log_T=22] [log_K=64]
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import zolt_b200 as Z  # noqa: E402
from zolt_b200 import host as H  # noqa: E402


def main():
    log_T = int(sys.argv[1]) if len(sys.argv) > 1 else 22
    log_K = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    ctx = Z.Context(0)
    rng = np.random.default_rng(9)
    idx = np.zeros((1 << log_T, 2), dtype=np.uint64)
    idx[:, 0] = rng.integers(0, 1 << 64, size=1 << log_T, dtype=np.uint64)
    if log_K < 64:
        idx[:, 0] &= np.uint64((1 << log_K) - 1)
    r = rng.integers(0, 1 << 64, size=(log_T, 4), dtype=np.uint64)
    r[:, 3] &= np.uint64((1 << 60) - 1)
    d = Z.DeviceBuffer.from_host(ctx, idx)
    ctx.timer_start()
    H.lasso_run(ctx, d.ptr, 1 << log_T, log_T, log_K, r)
    print(f"lasso T=2^{log_T} K=2^{log_K}: {ctx.timer_stop_ms():.3f} ms")
    d.free()


if __name__ == "__main__":
    main()
