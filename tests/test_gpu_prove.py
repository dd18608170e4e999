"""GPU parity: whole sumchecks driven by the C++ host mirror (Keccak transcript on the host, round
kernels on the GPU) against the CPU restatement running the same protocol.  Bit-exact: every round's
evaluations, every challenge, the final bound values."""
import numpy as np
import pytest

import zolt_b200 as Z
from oracle import cpu_ref as R
from oracle import poly as OP
from oracle import sumcheck as OS
from oracle import transcript as OT
from tests.util import P, fr_arr, fr_ints, rand_fr
from zolt_b200 import host as H

pytestmark = pytest.mark.gpu

KIND = {"sum": Z.SC_SUM, "prod2": Z.SC_PROD2, "prod3": Z.SC_PROD3, "eq_a_bmc": Z.SC_EQ_A_BMC, "eq_ab_minus_c": Z.SC_EQ_AB_MINUS_C}
NE = {"sum": 2, "prod2": 3, "prod3": 4, "eq_a_bmc": 4, "eq_ab_minus_c": 4}


def _rand_tables(seed, k, n):
    rng = np.random.default_rng(seed)
    t = rng.integers(0, 1 << 64, size=(k, n, 4), dtype=np.uint64)
    t[:, :, 3] &= np.uint64((1 << 60) - 1)
    return [t[i] for i in range(k)]


@pytest.mark.parametrize("kind,k", [("sum", 1), ("prod2", 2), ("prod3", 3), ("eq_a_bmc", 4), ("eq_ab_minus_c", 4)])
@pytest.mark.parametrize("pairing", [Z.PAIR_ADJACENT, Z.PAIR_HALF])
def test_prove_matches_cpu_reference(ctx, kind, k, pairing):
    log_n = 12
    tabs = _rand_tables(hash((kind, pairing)) % 1000, k, 1 << log_n)
    evals, chal, final = H.sumcheck_prove(ctx, KIND[kind], pairing, tabs)
    r_evals, r_chal, r_final = R.sc_prove(kind, pairing, tabs)
    ne = NE[kind]
    assert (chal == r_chal).all()
    assert (evals[:, :ne] == r_evals[:, :ne]).all()
    for t in range(k):
        assert (final[t] == r_final[t][0]).all()


def test_config2_deg2_eq_times_mle_2_20(ctx):
    # BASELINE config 2: Eq(tau, .) * M over 2^20 entries, adjacent pairs, host Keccak transcript
    log_n = 20
    tau = fr_arr(rand_fr(0x5A170020, log_n))
    eq_dev = Z.EqPolynomial.evals(ctx, tau)
    assert (eq_dev == R.eq_evals(tau)).all()
    m = _rand_tables(0x5A170021, 1, 1 << log_n)[0]
    evals, chal, final = H.sumcheck_prove(ctx, Z.SC_PROD2, Z.PAIR_ADJACENT, [eq_dev, m])
    r_evals, r_chal, r_final = R.sc_prove("prod2", 0, [eq_dev, m])
    assert (chal == r_chal).all() and (evals[:, :3] == r_evals[:, :3]).all()
    assert (final[0] == r_final[0][0]).all() and (final[1] == r_final[1][0]).all()
    # sumcheck invariant p(0) + p(1) = claim, claim_{j+1} = p_j(r_j)
    e0 = fr_ints(evals[0])
    claim = (e0[0] + e0[1]) % P
    for j in range(log_n):
        e = fr_ints(evals[j])
        assert (e[0] + e[1]) % P == claim
        claim = OP.evaluate_cubic_at_point(e, fr_ints(chal[j])[0])
    # final claim = Eq(r) * M(r) at the bound point
    assert claim == fr_ints(final[0])[0] * fr_ints(final[1])[0] % P


@pytest.mark.parametrize("with_c", [False, True])
def test_gruen_prove_matches_oracle(ctx, with_c):
    log_n = 10
    n = 1 << log_n
    k = 3 if with_c else 2
    tables = [rand_fr(40 + t, n) for t in range(k)]
    tau = rand_fr(50, log_n)
    eq = OP.eq_evals(tau)
    claim = sum(e * (a * b - (tables[2][i] if with_c else 0)) for i, (e, a, b) in enumerate(zip(eq, tables[0], tables[1]))) % P
    evals, chal, final_claim = H.sumcheck_prove_gruen(ctx, [fr_arr(t) for t in tables], fr_arr(tau), fr_arr([claim]), with_c)
    # oracle: same protocol with the Python restatements
    tr = OT.Transcript(b"bench")
    sq = OP.GruenSplitEq(tau)
    tabs = [list(t) for t in tables]
    cur = claim
    for rnd in range(log_n):
        t0, tinf = OS.round_gruen_ab(tabs, sq, with_c)
        s = sq.cubic_round_poly(t0, tinf, cur)
        assert fr_ints(evals[rnd]) == s, f"round {rnd}"
        assert (s[0] + s[1]) % P == cur
        for j, v in enumerate(s):
            tr.append_scalar(b"round_poly_%d" % j, v)
        r = tr.challenge_scalar(b"round")
        assert fr_ints(chal[rnd]) == [r]
        cur = OP.evaluate_cubic_at_point(s, r)
        tabs = OS.bind_tables(tabs, r, OS.ADJACENT)
        sq.bind(r)
    assert fr_ints(final_claim) == [cur]


def test_batched_sumcheck_prover_rlc(ctx):
    """S6: BatchedSumcheckProver (src/zkvm/batched_sumcheck.zig:77-430) over three GPU instances with different numbers of
    rounds (front-padded by scaled constant polynomials), caller-supplied batching coefficients: compressed round
    polynomials and the final claim match the oracle restatement."""
    from oracle import sumcheck as OS
    from tests.util import rand_fr

    inst = [("prod3", [rand_fr(300 + t, 1 << 9) for t in range(3)]),
            ("eq_ab_minus_c", [rand_fr(310 + t, 1 << 6) for t in range(4)]),
            ("sum", [rand_fr(320, 1 << 3)])]
    coeffs, challenges = rand_fr(330, 3), rand_fr(331, 9)
    want, want_final, claims = OS.run_batched(inst, coeffs, challenges)
    comb = {"prod3": Z.SC_PROD3, "eq_ab_minus_c": Z.SC_EQ_AB_MINUS_C, "sum": Z.SC_SUM}
    got, final = H.batched_sumcheck(ctx, [(comb[k], [fr_arr(t) for t in ts]) for k, ts in inst], fr_arr(claims), fr_arr(coeffs),
                                    fr_arr(challenges))
    assert [fr_ints(g) for g in got] == want
    assert fr_ints(final) == [want_final]


def test_dense_polynomial_evaluate_mirror(ctx):
    """DensePolynomial.evaluate (src/poly/mod.zig:73-92) through the C++ mirror: point[j] <-> variable j (index bit j)."""
    from oracle import bn254 as B
    from tests.util import P, fr_arr, fr_ints, rand_fr
    from zolt_b200 import host as H

    for nv in (0, 1, 5, 10):
        ev, pt = rand_fr(300 + nv, 1 << nv), rand_fr(400 + nv, nv)
        want = 0
        for i, v in enumerate(ev):
            w = v
            for j in range(nv):
                w = w * (pt[j] if (i >> j) & 1 else (1 - pt[j])) % P
            want = (want + w) % P
        assert fr_ints(H.dense_evaluate(ctx, fr_arr(ev), fr_arr(pt) if nv else [])) == [want]
