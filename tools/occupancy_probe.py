"""Fr-product throughput vs resident warps per SM (is the multiplier pipe latency- or issue-bound?)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import zolt_b200 as Z  # noqa: E402

ctx = Z.Context(0)
for kind, name in [(2, "fr_mul (2 chains/thread)"), (1, "imad_wide (8 chains/thread)")]:
    for threads, bps in [(32, 4), (64, 4), (128, 2), (128, 3), (128, 4), (256, 4), (256, 8)]:
        v = ctx.bench_int_pipe_cfg(kind, 400 if kind == 2 else 2000, threads, bps)
        print(f"{name:30s} warps/SM {threads * bps // 32:3d}: {v:.4g}/s", flush=True)
ctx.close()
