"""The Lasso oracle (oracle/lasso.py) replaying the reference's own unit tests: prover.zig "lasso prover claim tracking"
(:620-690: p(0) + p(1) = current_claim every round, the claim after receiveChallenge = p(challenge)), integration_test.zig
"lasso end-to-end simple case" (:68-120: the first round's g(0) + g(1) equals expanding_v.sum() = 1) and "lasso multiple rounds
consistent" (:282-338)."""
import json
import random
from pathlib import Path

import pytest

from oracle import lasso as L
from oracle.bn254 import FR_MOD as P
from oracle.poly import eq_evals


def _prover():
    # the reference tests' instance: 4 lookups {0,1,2,3}, log_T = 2, log_K = 3, r_reduction = (2, 3)
    return L.LassoProver([0, 1, 2, 3], 2, 3, [2, 3])


def test_reference_claim_tracking():
    pr = _prover()
    for rnd in range(5):
        claim = pr.current_claim
        c = pr.compute_round_polynomial()
        assert (c[0] + c[0] + c[1] + c[2]) % P == claim
        ch = rnd + 10
        pr.receive_challenge(ch)
        assert pr.current_claim == (c[0] + c[1] * ch + c[2] * ch * ch) % P
    assert pr.is_complete()


def test_reference_end_to_end_simple_case_and_multiple_rounds():
    pr = _prover()
    assert pr.current_claim == 1  # expanding_v.sum() of the fresh table: eq(r, .) sums to one over the full hypercube
    prev = 1
    for rnd in range(3):
        c = pr.compute_round_polynomial()
        assert (2 * c[0] + c[1] + c[2]) % P == prev
        ch = (rnd + 1) * 7
        prev = (c[0] + c[1] * ch + c[2] * ch * ch) % P
        pr.receive_challenge(ch)
    assert pr.round == 3 and not pr.is_address_phase()


@pytest.mark.parametrize("log_T,log_K,fill", [(1, 1, 2), (3, 5, 8), (4, 16, 11), (5, 12, 32)])
def test_table_convention_and_final_eval(log_T, log_K, fill):
    rng = random.Random(log_T * 100 + log_K)
    r = [rng.randrange(P) for _ in range(log_T)]
    idx = [rng.randrange(1 << log_K) for _ in range(fill)]
    pr = L.LassoProver(idx, log_T, log_K, r)
    # getEq (split_eq.zig:162-167) as one big-endian eq table: the point is (w_out reversed, w_in reversed)
    outer = log_T // 2
    be = r[:outer][::-1] + r[outer:][::-1]
    full = eq_evals(be)
    assert pr.eq_evals == [full[j] if j < fill else 0 for j in range(1 << log_T)]
    ch = [rng.randrange(P) for _ in range(log_T + log_K)]
    for c in ch:
        claim = pr.current_claim
        p = pr.compute_round_polynomial()
        assert (2 * p[0] + p[1]) % P == claim
        pr.receive_challenge(c)
    fe = 1
    for c in ch[:log_K]:
        fe = fe * (1 - c) % P
    assert pr.final_eval() == fe  # expanding_v.get(0)


def test_derive_challenge_is_a_u64_mix():
    a = L.derive_challenge([[1, 2, 3, 4], [0, 0, 0, 0], [5, 6, 7, 8]], 3)
    assert 0 <= a < 1 << 64 and a != L.derive_challenge([[1, 2, 3, 4], [0, 0, 0, 0], [5, 6, 7, 8]], 4)


def test_reference_run_round_tuples():
    """The 24 rounds LassoProver printed in the reference's captured run (logs/zolt.log, log_K = 16, log_T = 8; committed in
    tests/golden/zolt_log_kats.json): p(0) + p(1) is the claim, the claim the reference RE-SUMS from the updated table after
    receiveChallenge (prover.zig:402-407, :441-446) is p(challenge) -- which is how the device mirror obtains it without a
    pass over the table -- and the rounds chain.  Address rounds send degree-1 polynomials (c2 = 0).  The run's lookup indices
    are not in the log, so the tuples pin these identities, not the table values."""
    d = json.loads((Path(__file__).resolve().parent / "golden" / "zolt_log_kats.json").read_text())
    rounds = d["lasso_rounds"]
    assert len(rounds) == 24 and [r["round"] for r in rounds] == list(range(24))
    assert [r["phase"] for r in rounds] == ["address"] * 16 + ["cycle"] * 8
    rinv = pow(1 << 256, -1, P)
    std = lambda h: int(h, 16) * rinv % P  # noqa: E731  (printed limbs are Montgomery form)
    for i, r in enumerate(rounds):
        c0, c1, c2, ch = (std(r[k]) for k in ("c0", "c1", "c2", "challenge"))
        assert c2 == 0
        assert (2 * c0 + c1 + c2) % P == std(r["claim"])
        assert (c0 + c1 * ch + c2 * ch * ch) % P == std(r["claim_after"])
        if i + 1 < len(rounds):
            assert rounds[i + 1]["claim"] == r["claim_after"]
