"""LassoProver (src/zkvm/lasso/prover.zig) on the device against the loop-for-loop oracle (oracle/lasso.py): every round
polynomial's coefficients, the claim before and after every round, final_eval and -- in runLassoProver mode -- the challenges
the reference's deriveChallenge produces from them."""
import random

import numpy as np
import pytest

import zolt_b200 as Z
from oracle import bn254 as B
from oracle import lasso as OL
from oracle.bn254 import FR_MOD as P
from tests.util import fr_arr, fr_ints
from zolt_b200 import host as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built_lib):
    c = Z.Context(0)
    yield c
    c.close()


def _idx_array(idx):
    a = np.zeros((max(len(idx), 1), 2), dtype=np.uint64)
    for j, v in enumerate(idx):
        a[j, 0] = v & ((1 << 64) - 1)
        a[j, 1] = v >> 64
    return a


def _mont_limbs(x):
    return [int(v) for v in B.ints_to_mont_array([x], P)[0]]


def _check(ctx, idx, log_T, log_K, r, challenges):
    full = log_K <= 16
    d_idx = Z.DeviceBuffer.from_host(ctx, _idx_array(idx))
    try:
        coeffs, claims, chal, final = H.lasso_run(ctx, d_idx.ptr, len(idx), log_T, log_K, fr_arr(r) if log_T else None,
                                                  challenges=fr_arr(challenges) if challenges is not None else None)
    finally:
        d_idx.free()
    if challenges is None:
        exp_polys, exp_final, exp_ch = OL.run_lasso_prover(idx, log_T, log_K, r, _mont_limbs, full)
        assert fr_ints(chal) == exp_ch
        exp_claims = None
    else:
        pr = OL.LassoProver(idx, log_T, log_K, r, full)
        exp_polys, exp_claims = [], [pr.current_claim]
        for c in challenges:
            exp_polys.append(pr.compute_round_polynomial())
            pr.receive_challenge(c)
            exp_claims.append(pr.current_claim)
        exp_final = pr.final_eval()
    for rnd, e in enumerate(exp_polys):
        assert fr_ints(coeffs[rnd]) == e, f"round {rnd}"
    if exp_claims is not None:
        assert fr_ints(claims) == exp_claims
    assert fr_ints(final.reshape(1, 4)) == [exp_final]


def test_reference_unit_test_instance(ctx):
    # prover.zig "lasso prover claim tracking" (:620-690): challenges round + 10
    _check(ctx, [0, 1, 2, 3], 2, 3, [2, 3], [10, 11, 12, 13, 14])
    # integration_test.zig "lasso multiple rounds consistent" (:282-338): 7, 14, 21 then two more
    _check(ctx, [0, 1, 2, 3], 2, 3, [2, 3], [7, 14, 21, 28, 35])


@pytest.mark.parametrize("log_T,log_K,fill", [(0, 3, 1), (1, 1, 2), (2, 4, 3), (5, 8, 32), (7, 16, 100), (10, 64, 1000), (14, 128, 1 << 14),
                                              (15, 40, 30000)])
@pytest.mark.parametrize("mode", ["challenges", "derive"])
def test_lasso_prover_matches_oracle(ctx, log_T, log_K, fill, mode):
    rng = random.Random(log_T * 1000 + log_K)
    r = [rng.randrange(P) for _ in range(log_T)]
    idx = [rng.randrange(1 << log_K) for _ in range(fill)]
    ch = [rng.randrange(P) for _ in range(log_T + log_K)] if mode == "challenges" else None
    _check(ctx, idx, log_T, log_K, r, ch)


def test_addr_round_primitive(ctx):
    """zolt_lasso_addr_round alone: bind by one bit, sum by another, sum_bit < 0 = the total"""
    rng = random.Random(5)
    n = 3000
    eq = [rng.randrange(P) for _ in range(n)]
    idx = [rng.randrange(1 << 128) for _ in range(n)]
    r = rng.randrange(P)
    d_eq = Z.DeviceBuffer.from_host(ctx, fr_arr(eq))
    d_idx = Z.DeviceBuffer.from_host(ctx, _idx_array(idx))
    rr = fr_arr([r])
    try:
        for bind_bit, sum_bit in ((-1, 0), (0, 97), (97, 127), (127, -1)):
            sums = ctx.lasso_addr_round(d_eq.ptr, d_idx.ptr, n, bind_bit, rr, sum_bit)
            if bind_bit >= 0:
                eq = [e * (r if (i >> bind_bit) & 1 else (1 - r) % P) % P for e, i in zip(eq, idx)]
            exp = [0, 0]
            for e, i in zip(eq, idx):
                b = (i >> sum_bit) & 1 if sum_bit >= 0 else 0
                exp[b] = (exp[b] + e) % P
            assert fr_ints(sums) == exp
        assert fr_ints(d_eq.download((n, 4))) == eq
    finally:
        d_eq.free()
        d_idx.free()
