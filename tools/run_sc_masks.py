import sys, numpy as np
sys.path.insert(0, '/root/repo')
import zolt_b200 as Z
from zolt_b200 import host as H
ctx = Z.Context(0)
rng = np.random.default_rng(3)
n = 1 << 24
bufs = []
for _ in range(4):
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 60) - 1)
    bufs.append(Z.DeviceBuffer.from_host(ctx, a))
for mask in [int(x, 0) for x in sys.argv[1].split(',')]:
    ctx.set_tuning(2, mask)
    inst = Z.SumcheckInstance.from_device(ctx, Z.SC_EQ_AB_MINUS_C, Z.PAIR_ADJACENT, [b.ptr for b in bufs], 24)
    ch = rng.integers(0, 1 << 64, size=(2, 4), dtype=np.uint64); ch[:, 3] &= np.uint64((1 << 60) - 1)
    for j in range(2):
        inst.compute_round_poly(); inst.bind_challenge(ch[j])
    ctx.sync(); inst.free()
ctx.close()
