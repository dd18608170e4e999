"""Per-phase device timings (CUDA events inside the library) for one MSM and one sumcheck."""
import json
import os
import sys
from collections import defaultdict
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import zolt_b200 as Z  # noqa: E402

TAGS = {1: "digits", 2: "scan", 3: "scatter", 4: "accum0", 5: "edges", 6: "reduce", 7: "affine", 16: "sc_eval", 17: "sc_fused"}


def rand_mont(rng, n):
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


def main():
    ctx = Z.Context(0)
    rng = np.random.default_rng(1)
    out = {}
    for log_n in [int(x) for x in os.environ.get("PHASE_MSM", "20,24").split(",") if x]:
        n = 1 << log_n
        bases = Z.Bases.fixed_base_mul(ctx, rand_mont(rng, n))
        d_s = Z.DeviceBuffer.from_host(ctx, rand_mont(rng, n))
        Z.MSM.compute_dev(bases, d_s.ptr, n)  # warm-up
        ctx.prof_enable(True)
        ctx.timer_start()
        Z.MSM.compute_dev(bases, d_s.ptr, n)
        total = ctx.timer_stop_ms()
        ph = {TAGS[t]: round(ms, 4) for t, ms in ctx.prof_read()}
        ctx.prof_enable(False)
        out[f"msm_2^{log_n}"] = {"total_ms": round(total, 3), **ph}
        print(f"msm 2^{log_n}: total {total:.3f} ms", ph, flush=True)
        for c in [int(x) for x in os.environ.get("PHASE_PRE", "22").split(",") if x]:
            import time
            t0 = time.time()
            bases.precompute(c)
            t_pre = time.time() - t0
            r0 = Z.MSM.compute_dev(bases, d_s.ptr, n)  # warm-up (workspace growth)
            ctx.prof_enable(True)
            ctx.timer_start()
            r1 = Z.MSM.compute_dev(bases, d_s.ptr, n)
            total = ctx.timer_stop_ms()
            ph = {TAGS[t]: round(ms, 4) for t, ms in ctx.prof_read()}
            ctx.prof_enable(False)
            out[f"msm_2^{log_n}_pre{c}"] = {"total_ms": round(total, 3), "precompute_s": round(t_pre, 3), **ph}
            print(f"msm 2^{log_n} table c={c}: total {total:.3f} ms (table built in {t_pre:.2f} s)", ph, flush=True)
        d_s.free()
        bases.free()
    for name, comb, k, log_n in [("eq_ab_minus_c", Z.SC_EQ_AB_MINUS_C, 4, 24), ("prod2", Z.SC_PROD2, 2, 20)]:
        n = 1 << log_n
        tabs = [rand_mont(rng, n) for _ in range(k)]
        chal = rand_mont(rng, log_n)
        for rep in range(2):
            inst = Z.SumcheckInstance(ctx, comb, Z.PAIR_ADJACENT, tabs)
            ctx.prof_enable(rep == 1)
            ctx.timer_start()
            for j in range(log_n):
                inst.compute_round_poly()
                inst.bind_challenge(chal[j])
            total = ctx.timer_stop_ms()
            if rep == 1:
                rounds = [round(ms, 4) for _, ms in ctx.prof_read()]
                ctx.prof_enable(False)
                out[f"sc_{name}_2^{log_n}"] = {"total_ms": round(total, 3), "kernel_sum_ms": round(sum(rounds), 3), "rounds_ms": rounds}
                print(f"sumcheck {name} 2^{log_n}: wall {total:.3f} ms, kernels {sum(rounds):.3f} ms", rounds, flush=True)
            inst.free()
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "phases.json").write_text(json.dumps(out, indent=1))
    ctx.close()


if __name__ == "__main__":
    main()
