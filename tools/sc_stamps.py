"""Per-launch timeline of the sumcheck round kernels (globaltimer stamps, tuning key 7); diagnostics, not a benchmark.

    python tools/sc_stamps.py [log_n=20] [combiner=eq_ab_minus_c]
Columns (us, relative to the kernel's first stamp): loop = block 0 leaves its streaming loop, last = the last block has its ticket,
fold = partials folded, fin = round polynomial computed, end = mailbox written; `prof` = the launch's event-timed duration.
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import zolt_b200 as Z  # noqa: E402
from zolt_b200 import host as H  # noqa: E402

COMB = {"prod2": (Z.SC_PROD2, 2), "prod3": (Z.SC_PROD3, 3), "eq_ab_minus_c": (Z.SC_EQ_AB_MINUS_C, 4)}


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    name = sys.argv[2] if len(sys.argv) > 2 else "eq_ab_minus_c"
    comb, k = COMB[name]
    ctx = Z.Context(0)
    rng = np.random.default_rng(3)
    n = 1 << log_n
    bufs = []
    for _ in range(k):
        a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
        a[:, 3] &= np.uint64((1 << 60) - 1)
        bufs.append(Z.DeviceBuffer.from_host(ctx, a))
    ptrs = [b.ptr for b in bufs]
    for key in ("SC_SMALL", "SC_TAIL"):
        if os.environ.get(key) is not None:
            ctx.set_tuning(1 if key == "SC_SMALL" else 4, int(os.environ[key]))
    for _ in range(2):
        H.sumcheck_prove_dev(ctx, comb, Z.PAIR_ADJACENT, ptrs, log_n)
    dbg = Z.DeviceBuffer(ctx, 8 * 8 * 64)
    dbg.upload(np.zeros(8 * 64, dtype=np.uint64))
    ctx.set_tuning(7, dbg.ptr)
    ctx.prof_enable(True)
    ctx.timer_start()
    H.sumcheck_prove_dev(ctx, comb, Z.PAIR_ADJACENT, ptrs, log_n)
    wall = ctx.timer_stop_ms()
    log = ctx.prof_read()
    ctx.prof_enable(False)
    ctx.sync()
    ctx.set_tuning(7, 0)
    t = dbg.download((64, 8)).astype(np.int64)
    rows = [r for r in t if r[1] or r[5]]
    rows.sort(key=lambda r: r[5])
    ks = [(tag, ms) for tag, ms in log if tag in (16, 17, 18)]
    print(f"log_n {log_n} {name}: wall {wall:.3f} ms, kernels {sum(m for _, m in ks):.3f} ms, {len(ks)} launches")
    print("round  prof_us |  loop   last   fold    fin    end | gap to next start")
    for i, r in enumerate(rows):
        base = next((x for x in r[:6] if x), 0)
        rel = [(x - base) / 1e3 if x else float('nan') for x in r[1:6]]
        nxt = rows[i + 1][0] if i + 1 < len(rows) and rows[i + 1][0] else 0
        gap = (nxt - r[5]) / 1e3 if nxt else float('nan')
        prof = ks[i][1] * 1e3 if i < len(ks) else float('nan')
        print(f"{i:4d} {prof:8.1f} | " + " ".join(f"{x:6.1f}" for x in rel) + f" | {gap:6.1f}")


if __name__ == "__main__":
    main()
