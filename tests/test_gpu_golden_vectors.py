"""GPU parity against COMMITTED vectors (tests/golden/oracle_vectors.json, written by tests/golden/make_oracle_vectors.py
from the pinned oracle): MSM, HyperKZG commit / open, every round of the dense combiners and the Gruen sums.  Bit-exact."""
import json
from pathlib import Path

import numpy as np
import pytest

import zolt_b200 as Z
from oracle import bn254 as B
from tests.util import fr_arr, fr_ints

pytestmark = pytest.mark.gpu
V = json.loads((Path(__file__).resolve().parent / "golden" / "oracle_vectors.json").read_text())
COMB = {"eq_ab_minus_c": Z.SC_EQ_AB_MINUS_C, "prod3": Z.SC_PROD3, "prod2": Z.SC_PROD2, "eq_a_bmc": Z.SC_EQ_A_BMC, "sum": Z.SC_SUM}


def _ints(xs):
    return [int(x, 16) for x in xs]


def _pt(p):
    return None if p is None else (int(p[0], 16), int(p[1], 16))


def test_msm_vector(ctx):
    logs, scalars = _ints(V["msm"]["base_logs"]), _ints(V["msm"]["scalars"])
    bases = Z.Bases.fixed_base_mul(ctx, fr_arr(logs))
    xy, inf = Z.MSM.compute(bases, fr_arr(scalars))
    assert B.arrays_to_points(xy.reshape(1, 8), [inf])[0] == _pt(V["msm"]["result"])
    bases.free()


def test_hyperkzg_vectors(ctx):
    hk = V["hyperkzg"]
    srs = Z.HyperKZG.setup(ctx, 8)
    xy, inf = Z.HyperKZG.commit(srs, fr_arr(_ints(hk["evals"])))
    assert B.arrays_to_points(xy.reshape(1, 8), [inf])[0] == _pt(hk["commitment"])
    qxy, qinf, fin = Z.HyperKZG.open(srs, fr_arr(_ints(hk["evals"])), fr_arr(_ints(hk["point"])))
    assert B.arrays_to_points(qxy, qinf) == [_pt(p) for p in hk["quotients"]]
    assert fr_ints(fin) == [int(hk["final_eval"], 16)]
    srs.free()


@pytest.mark.parametrize("kind", sorted(V["sumcheck"]))
@pytest.mark.parametrize("small_pairs", [1 << 14, 0])
def test_sumcheck_vectors(ctx, kind, small_pairs):
    s = V["sumcheck"][kind]
    ctx.set_tuning(Z.TUNE_SC_SMALL_PAIRS, small_pairs)  # small-round kernel / streaming kernels
    try:
        inst = Z.SumcheckInstance(ctx, COMB[kind], s["pairing"], [fr_arr(_ints(t)) for t in s["tables"]])
        for want, r in zip(s["round_evals"], _ints(s["challenges"])):
            assert fr_ints(inst.compute_round_poly()) == _ints(want)
            inst.bind_challenge(fr_arr([r]))
        assert [fr_ints(inst.table(t))[0] for t in range(len(s["tables"]))] == _ints(s["final_values"])
        inst.free()
    finally:
        ctx.set_tuning(Z.TUNE_SC_SMALL_PAIRS, 1 << 14)


def test_gruen_vector(ctx):
    g = V["gruen_ab"]
    inst = Z.SumcheckInstance(ctx, Z.SC_GRUEN_AB, Z.PAIR_ADJACENT, [fr_arr(_ints(t)) for t in g["tables"]], tau=fr_arr(_ints(g["tau"])))
    for want, r in zip(g["round_sums"], _ints(g["challenges"])):
        assert fr_ints(inst.compute_round_poly()) == _ints(want)
        inst.bind_challenge(fr_arr([r]))
    inst.free()
