"""MSM 2^24 over window-level tables of different window sizes (diagnostics): c = 20 is the default; 22 saves a window
(accumulation 31.8 -> 29.8 ms) and loses it again in the bucket reduction (4x the buckets) and the bin partition.

    python tools/time_msm_window.py
"""
import sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np
import zolt_b200 as Z
n = 1 << 24
ctx = Z.Context(0)
rng = np.random.default_rng(5)
k = rng.integers(1, 1 << 62, size=(n, 4), dtype=np.uint64); k[:, 3] &= np.uint64((1 << 59) - 1)
bases = Z.Bases.fixed_base_mul(ctx, k)
s = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64); s[:, 3] &= np.uint64((1 << 60) - 1)
d_s = Z.DeviceBuffer.from_host(ctx, s)
TAGS = {1: "part", 2: "count", 3: "place", 4: "accum0", 5: "edges", 6: "reduce", 7: "affine"}
ref = None
for c in (20, 22, 21, 23):
    bases.precompute(c)
    Z.MSM.compute_dev(bases, d_s.ptr, n)
    best = None
    for _ in range(3):
        ctx.prof_enable(True); ctx.timer_start()
        out = Z.MSM.compute_dev(bases, d_s.ptr, n)
        wall = ctx.timer_stop_ms(); log = ctx.prof_read(); ctx.prof_enable(False)
        if best is None or wall < best[0]: best = (wall, log)
    ph = {}
    for tag, ms in best[1]: ph[TAGS.get(tag, tag)] = round(ph.get(TAGS.get(tag, tag), 0) + ms, 3)
    if ref is None: ref = out
    print("c", c, f"{best[0]:.3f} ms", ph, "same:", bool((out[0] == ref[0]).all()), flush=True)
