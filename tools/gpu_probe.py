"""First-contact probe for a B200 box: integer-pipe micro-benchmarks and coarse timings of the two
hot paths at scale.  Writes gpurun_out/probe.json.  Not a benchmark of record (see bench.py)."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import zolt_b200 as Z  # noqa: E402


def rand_mont(rng, n):
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)  # < 2^252 < r: canonical limbs
    return a


def main():
    out = {}
    ctx = Z.Context(0)
    names = {0: "imad_reg", 4: "imad_imm", 1: "imad_wide", 2: "fr_mul", 3: "fq_mul", 5: "dfma", 6: "dfma_with_imad"}
    for kind, name in names.items():
        out[name + "_per_s"] = ctx.bench_int_pipe(kind, 4000 if kind in (0, 1, 4, 5, 6) else 400)
        print(name, f"{out[name + '_per_s']:.4g}", flush=True)
    rng = np.random.default_rng(1)
    for log_n in [int(x) for x in os.environ.get("PROBE_MSM", "16,20,22,24").split(",") if x]:
        n = 1 << log_n
        t0 = time.time()
        bases = Z.Bases.fixed_base_mul(ctx, rand_mont(rng, n))
        t_gen = time.time() - t0
        sc = rand_mont(rng, n)
        d_s = Z.DeviceBuffer.from_host(ctx, sc)
        best = 1e9
        for rep in range(3):
            ctx.timer_start()
            xy, inf = Z.MSM.compute_dev(bases, d_s.ptr, n)
            best = min(best, ctx.timer_stop_ms())
        out[f"msm_2^{log_n}_ms"] = best
        out[f"fixed_base_2^{log_n}_s"] = t_gen
        print(f"msm 2^{log_n}: {best:.2f} ms (bases generated in {t_gen:.2f} s) inf={inf}", flush=True)
        d_s.free()
        bases.free()
    for name, comb, k, log_n in [("prod2", Z.SC_PROD2, 2, 20), ("eq_ab_minus_c", Z.SC_EQ_AB_MINUS_C, 4, 24),
                                 ("prod3", Z.SC_PROD3, 3, 24), ("sum", Z.SC_SUM, 1, 24)]:
        if os.environ.get("PROBE_SC", "1") == "0":
            break
        n = 1 << log_n
        tabs = [rand_mont(rng, n) for _ in range(k)]
        chal = rand_mont(rng, log_n)
        inst = Z.SumcheckInstance(ctx, comb, Z.PAIR_ADJACENT, tabs)
        ctx.timer_start()
        for j in range(log_n):
            inst.compute_round_poly()
            inst.bind_challenge(chal[j])
        ms = ctx.timer_stop_ms()
        gb = 128 * k * n / 1e9
        out[f"sc_{name}_2^{log_n}_ms"] = ms
        out[f"sc_{name}_2^{log_n}_GBps"] = gb / (ms * 1e-3)
        print(f"sumcheck {name} 2^{log_n}: {ms:.3f} ms  {gb / (ms * 1e-3):.1f} GB/s algorithmic", flush=True)
        inst.free()
    out["launches"] = ctx.launch_count
    os.makedirs(ROOT / "gpurun_out", exist_ok=True)
    (ROOT / "gpurun_out" / "probe.json").write_text(json.dumps(out, indent=1))
    ctx.close()


if __name__ == "__main__":
    main()
