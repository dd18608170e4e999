"""The Stage 4 oracle (oracle/stage4.py) against the protocol's own invariants: every round polynomial satisfies
p(0) + p(1) = claim, the last claim is eq(r_cycle, r) * C(bound values), and the merged eq equals the closed form; plus the
reference's own unit expectations for the split eq (gruen_eq.zig:343-420)."""
import random

import pytest

from oracle import stage4 as S
from oracle.bn254 import FR_MOD as P
from oracle.poly import eq_mle


@pytest.mark.parametrize("log_T", [1, 2, 3, 4])
def test_stage4_oracle_invariants(log_T):
    rng = random.Random(log_T)
    T = 1 << log_T
    n = S.K * T
    tabs = [[rng.randrange(P) for _ in range(n)] for _ in range(4)]
    inc = [rng.randrange(P) for _ in range(T)]
    rc = [rng.randrange(P) for _ in range(log_T)]
    for p1 in range(1, log_T + 1):
        pr = S.Stage4Prover(*tabs, inc, rng.randrange(P), rc, p1, S.LOG_K)
        ch = [rng.randrange(P) for _ in range(pr.num_rounds)]
        claim0 = pr.compute_input_claim()
        evals, final = pr.run(ch, claim0)
        c = claim0
        for e, r in zip(evals, ch):
            assert (e[0] + e[1]) % P == c
            c = S.eval_coeffs(S.coeffs_from_evals(e), r)
        fc = pr.final_claims()
        ra = (pr.gamma * fc["rs1_ra"] + pr.gamma_sq * fc["rs2_ra"]) % P
        body = (ra * fc["val"] + fc["rd_wa"] * (fc["val"] + fc["inc"])) % P
        assert pr.merged_eq[0] * body % P == final
        # cycle challenges: rounds 0..p1-1 and the last log_T - p1; round order is little-endian, r_cycle too
        cyc = ch[:p1] + ch[p1 + S.LOG_K:]
        assert pr.merged_eq[0] == eq_mle(rc, cyc)


def test_gruen_eq_reference_unit_expectations():
    # gruen_eq.zig "gruen eq polynomial initialization" (:337-356)
    g = S.GruenEq([1, 2, 3, 4])
    assert g.current_index == 4 and g.current_scalar == 1
    assert len(g.E_out_current()) == 4 and len(g.E_in_current()) == 2  # m = 2 out bits, n - 1 - m = 1 in bit
    # "gruen eq polynomial bind" (:358-385): w = [2, 3], r = 5 -> index 1, scalar = eq(3, 5) = 1 - 3 - 5 + 30
    g = S.GruenEq([2, 3])
    g.bind(5)
    assert g.current_index == 1 and g.current_scalar == (1 - 3 - 5 + 30) % P
    m = g.merge()  # eq(w[0..1], .) scaled (:172-203)
    assert len(m) == 2 and sum(m) % P == g.current_scalar
    # "gruen poly deg 3" (:387-414): p(0) + p(1) = claim
    c = S.GruenEq([1, 2]).gruen_poly_deg3(10, 5, 100)
    assert (c[0] + sum(c)) % P == 100


def test_gruen_eq_tables_match_the_reference_run():
    """Value-level vectors from the reference's own captured run (logs/zolt.log -> tests/golden/zolt_log_kats.json): Stage 4's
    r_cycle is the Stage 3 challenge vector (round order; the split eq takes it big-endian, stage4_gruen_prover.zig:259-281),
    and STAGE4_GRUEN_INIT printed E_out.len = 16, E_in.len = 8 and the first four entries of each table
    (gruen_eq.zig evalsCached :296-334).  The first two STAGE4_CONTRIB lines give E_out[x_out] * E_in[x_in] (low 8 bytes)."""
    import json
    from pathlib import Path

    d = json.loads((Path(__file__).resolve().parent / "golden" / "zolt_log_kats.json").read_text())
    le = lambda b: int.from_bytes(bytes(b), "little")  # noqa: E731
    ch = [le(k["challenge"]) for k in sorted((k for k in d["kats"] if k["stage"] == "STAGE3"), key=lambda x: x["round"])]
    assert len(ch) == 8
    g = S.GruenEq(ch[::-1])
    e_out, e_in = g.E_out_current(), g.E_in_current()
    assert len(e_out) == 16 and len(e_in) == 8
    init = d["stage4_gruen_init"]
    assert e_out[:4] == [le(v) for v in init["E_out"]]
    assert e_in[:4] == [le(v) for v in init["E_in"]]
    # [STAGE4_CONTRIB #1] x_out=0, x_in=0 and [#2] x_out=0, x_in=1: E_combined, first 8 bytes
    assert (e_out[0] * e_in[0] % P).to_bytes(32, "little")[:8] == bytes([35, 129, 211, 100, 28, 19, 13, 225])
    assert (e_out[0] * e_in[1] % P).to_bytes(32, "little")[:8] == bytes([102, 202, 125, 226, 26, 250, 145, 82])


def test_gruen_poly_deg3_and_bind_match_the_reference_run():
    """The first phase-1 rounds of the reference's captured Stage 4 run: from the logged weighted sums q(0), q_X2 and the claim,
    gruenPolyDeg3 (gruen_eq.zig:214-262) must return the logged cubic c0..c3; the second round only does after bind
    (:120-160) has taken the first round's challenge, so the running scalar and the bound-variable bookkeeping are pinned too."""
    import json
    from pathlib import Path

    d = json.loads((Path(__file__).resolve().parent / "golden" / "zolt_log_kats.json").read_text())
    le = lambda b: int.from_bytes(bytes(b), "little")  # noqa: E731
    ch = [le(k["challenge"]) for k in sorted((k for k in d["kats"] if k["stage"] == "STAGE3"), key=lambda x: x["round"])]
    g = S.GruenEq(ch[::-1])
    rounds = d["stage4_gruen_init"]["phase1_rounds"]
    assert [r["round"] for r in rounds] == [0, 1]
    for r in rounds:
        got = g.gruen_poly_deg3(le(r["q_0"]), le(r["q_X2"]), le(r["previous_claim"]))
        assert list(got) == [le(r[f"c{i}"]) for i in range(4)], r["round"]
        assert (2 * got[0] + got[1] + got[2] + got[3]) % P == le(r["previous_claim"])
        g.bind(le(r["challenge"]))
