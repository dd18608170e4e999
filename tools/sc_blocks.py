"""When does each block of one streaming round kernel leave its loop, and on which SM?  (diagnostics, tuning key 7)

    python tools/sc_blocks.py [log_n=24] [fused=0|1]
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import zolt_b200 as Z  # noqa: E402


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    fused = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    ctx = Z.Context(0)
    rng = np.random.default_rng(3)
    n = 1 << log_n
    bufs = []
    for _ in range(4):
        a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
        a[:, 3] &= np.uint64((1 << 60) - 1)
        bufs.append(Z.DeviceBuffer.from_host(ctx, a))
    dbg = Z.DeviceBuffer(ctx, 8 * (512 + 2 * 1024))
    for rep in range(3):
        sc = Z.SumcheckInstance.from_device(ctx, Z.SC_EQ_AB_MINUS_C, Z.PAIR_ADJACENT, [b.ptr for b in bufs], log_n)
        dbg.upload(np.zeros(512 + 2048, dtype=np.uint64))
        ctx.set_tuning(7, dbg.ptr)
        sc.compute_round_poly()
        if fused:
            ctx.sync()
            dbg.upload(np.zeros(512 + 2048, dtype=np.uint64))
            sc.bind_challenge(np.array([5, 6, 7, 8], dtype=np.uint64))
            sc.compute_round_poly()
        ctx.sync()
        ctx.set_tuning(7, 0)
        t = dbg.download((512 + 2048,)).astype(np.int64)
        sc.free()
    head = t[:512].reshape(64, 8)
    row = head[np.argmax(head[:, 5])]
    base = row[0] if row[0] else None
    blk = t[512:].reshape(-1, 2)
    blk = blk[blk[:, 0] > 0]
    ends = blk[:, 0]
    t0 = base if base else ends.min()
    rel = (ends - t0) / 1e3
    print(f"log_n {log_n} fused {fused}: {len(blk)} blocks; kernel end stamp {(row[5] - t0) / 1e3:.1f} us; block loop-end us: min {rel.min():.1f} "
          f"p10 {np.percentile(rel, 10):.1f} p50 {np.percentile(rel, 50):.1f} p90 {np.percentile(rel, 90):.1f} max {rel.max():.1f}")
    hist, edges = np.histogram(rel, bins=12)
    for h, e0, e1 in zip(hist, edges[:-1], edges[1:]):
        print(f"  {e0:8.1f} .. {e1:8.1f} us: {h:4d} " + "#" * (h // 4))
    # per SM: spread of the end times of its blocks
    sms = {}
    for (e, sm), r in zip(blk, rel):
        sms.setdefault(int(sm), []).append(float(r))
    per = sorted((max(v), min(v), k, len(v)) for k, v in sms.items())
    print("  SMs:", len(sms), "blocks/SM:", sorted(set(len(v) for v in sms.values())))
    print("  fastest SMs (last block end, first block end, smid, n):", [(round(a, 1), round(b, 1), k, c) for a, b, k, c in per[:4]])
    print("  slowest SMs:", [(round(a, 1), round(b, 1), k, c) for a, b, k, c in per[-4:]])
    order = np.argsort(blk[:, 0])
    print("  first blocks to finish (block idx):", np.nonzero(t[512::2] > 0)[0][order[:8]].tolist(), " last:", np.nonzero(t[512::2] > 0)[0][order[-8:]].tolist())


if __name__ == "__main__":
    main()
