"""Run the first rounds of one sumcheck (for `ncu -k regex:sc_round`); not a benchmark.

    python tools/run_sc_once.py [combiner=eq_ab_minus_c] [log_n=24] [rounds=3]
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import zolt_b200 as Z  # noqa: E402

COMB = {"sum": (Z.SC_SUM, 1), "prod2": (Z.SC_PROD2, 2), "prod3": (Z.SC_PROD3, 3), "eq_a_bmc": (Z.SC_EQ_A_BMC, 4),
        "eq_ab_minus_c": (Z.SC_EQ_AB_MINUS_C, 4)}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "eq_ab_minus_c"
    log_n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    comb, k = COMB[name]
    ctx = Z.Context(0)
    rng = np.random.default_rng(3)
    n = 1 << log_n
    tabs = []
    for _ in range(k):
        a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
        a[:, 3] &= np.uint64((1 << 60) - 1)
        tabs.append(a)
    chal = rng.integers(0, 1 << 64, size=(rounds, 4), dtype=np.uint64)
    chal[:, 3] &= np.uint64((1 << 60) - 1)
    inst = Z.SumcheckInstance(ctx, comb, Z.PAIR_ADJACENT, tabs)
    for j in range(rounds):
        inst.compute_round_poly()
        inst.bind_challenge(chal[j])
    ctx.sync() if hasattr(ctx, "sync") else None
    inst.free()
    ctx.close()


if __name__ == "__main__":
    main()
