"""Per-warp timeline of the MSM level-0 accumulation kernel (tuning key 8); diagnostics.

    python tools/msm_blocks.py [log_n=24]
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import zolt_b200 as Z  # noqa: E402


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
    d_s = Z.DeviceBuffer.from_host(ctx, s)
    Z.MSM.compute_dev(bases, d_s.ptr, n)
    words = 3 * (13 * n // 64 // 32 + 4096) + 64
    dbg = Z.DeviceBuffer(ctx, 8 * words)
    dbg.upload(np.zeros(words, dtype=np.uint64))
    ctx.set_tuning(8, dbg.ptr)
    ctx.prof_enable(True)
    Z.MSM.compute_dev(bases, d_s.ptr, n)
    log = ctx.prof_read()
    ctx.prof_enable(False)
    ctx.sync()
    ctx.set_tuning(8, 0)
    t = dbg.download((words,)).astype(np.int64)[: 3 * ((words - 64) // 3)].reshape(-1, 3)
    t = t[t[:, 1] > 0]
    t0 = t[:, 0].min()
    st, en = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3
    print("prof:", [(tag, round(ms, 3)) for tag, ms in log])
    print(f"log_n {log_n}: {len(t)} warps ({len(t) // 4} blocks); kernel span {en.max():.0f} us; warp duration us: min {np.min(en - st):.0f} p10 {np.percentile(en - st, 10):.0f} "
          f"p50 {np.percentile(en - st, 50):.0f} p90 {np.percentile(en - st, 90):.0f} max {np.max(en - st):.0f}")
    # resident warps over time
    T = en.max()
    edges = np.linspace(0, T, 41)
    act = [(np.sum((st <= x) & (en > x))) for x in edges[:-1]]
    print("  resident warps at 40 sample times:", act)
    # when did the last block START, and how many warps were still running in the last 10 % / 5 % / 2 % of the span
    print(f"  last warp start {st.max():.0f} us; running at 90% of span: {np.sum(en > 0.9 * T)}, at 95%: {np.sum(en > 0.95 * T)}, at 98%: {np.sum(en > 0.98 * T)} of {148 * 16} slots")


if __name__ == "__main__":
    main()
