"""CPU tier: the C++ restatement (oracle/cpu_ref.cpp, the timed CPU baseline) against the Python
big-integer oracle.  Bit-exact."""
import numpy as np
import pytest

from oracle import bn254 as B
from oracle import cpu_ref as R
from oracle import poly as OP
from oracle import sumcheck as OS
from oracle import transcript as OT
from tests.util import P, fr_arr, fr_ints, rand_fr


@pytest.mark.parametrize("field,mod", [(0, B.FR_MOD), (1, B.FQ_MOD)])
def test_field_ops(field, mod):
    rng = np.random.default_rng(3)
    n = 2000
    av = [v % mod for v in B.random_fr(rng, n)]
    bv = [v % mod for v in B.random_fr(rng, n)]
    av[:4], bv[:4] = [0, 1, mod - 1, mod - 1], [0, mod - 1, mod - 1, 1]
    a, b = B.ints_to_raw_array(av), B.ints_to_raw_array(bv)
    rinv = pow(1 << 256, -1, mod)
    assert B.raw_array_to_ints(R.field_op(field, 0, a, b)) == [x * y * rinv % mod for x, y in zip(av, bv)]
    assert B.raw_array_to_ints(R.field_op(field, 1, a, b)) == [(x + y) % mod for x, y in zip(av, bv)]
    assert B.raw_array_to_ints(R.field_op(field, 2, a, b)) == [(x - y) % mod for x, y in zip(av, bv)]
    assert B.raw_array_to_ints(R.field_op(field, 6, a)) == [(-x) % mod for x in av]
    Rm = (1 << 256) % mod
    got = B.raw_array_to_ints(R.field_op(field, 3, a[:64]))
    assert got == [0 if x == 0 else pow(x * rinv % mod, -1, mod) * Rm % mod for x in av[:64]]


@pytest.mark.parametrize("n", [0, 1, 5, 8, 40, 600])
def test_msm_matches_python(n):
    logs = [k or 1 for k in rand_fr(300 + n, n)]
    scalars = rand_fr(400 + n, n)
    xy, inf = R.g1_fixed_mul(fr_arr(logs) if n else np.empty((0, 4), dtype=np.uint64))
    if n <= 40:
        assert B.arrays_to_points(xy, inf) == [B.g1_mul(B.G1_GEN, k) for k in logs]
    out, oinf = R.msm(xy, inf, fr_arr(scalars) if n else np.empty((0, 4), dtype=np.uint64))
    got = B.arrays_to_points(out.reshape(1, 8), [oinf])[0]
    assert got == B.msm_closed_form(logs, scalars)
    if got is None:
        assert oinf == 1 and not out.any()


def test_parallel_msm_equals_sequential():
    # src/msm/mod.zig:911-936 (Parallel == Sequential), here on 4096 points so chunks are >= 1024
    n = 4096
    logs, scalars = rand_fr(1, n), rand_fr(2, n)
    xy, inf = R.g1_fixed_mul(fr_arr(logs), threads=8)
    a = R.msm(xy, inf, fr_arr(scalars), threads=1)
    b = R.msm(xy, inf, fr_arr(scalars), threads=4)
    assert (a[0] == b[0]).all() and a[1] == b[1]
    assert B.arrays_to_points(a[0].reshape(1, 8), [a[1]])[0] == B.msm_closed_form(logs, scalars)


def test_poly_ops():
    t = rand_fr(7, 64)
    r = rand_fr(8, 1)
    assert fr_ints(R.bind_low(fr_arr(t), fr_arr(r))) == OP.bind_low(t, r[0])
    assert fr_ints(R.bind_first(fr_arr(t), fr_arr(r))) == OP.bind_first(t, r[0])
    tau = rand_fr(9, 7)
    assert fr_ints(R.eq_evals(fr_arr(tau))) == OP.eq_evals(tau)
    assert fr_ints(R.eq_evals(fr_arr(tau), fr_arr(r))) == OP.eq_evals(tau, r[0])


@pytest.mark.parametrize("kind,k", [("sum", 1), ("prod2", 2), ("prod3", 3), ("eq_a_bmc", 4), ("eq_ab_minus_c", 4)])
@pytest.mark.parametrize("pairing", [0, 1])
def test_round_loops(kind, k, pairing):
    tabs = [rand_fr(50 + i, 128) for i in range(k)]
    got = fr_ints(R.sc_round(kind, pairing, [fr_arr(t) for t in tabs]))
    assert got == OS.ROUND_FNS[kind](tabs, pairing)


@pytest.mark.parametrize("with_c", [False, True])
def test_gruen_round(with_c):
    n_vars = 7
    tau = rand_fr(60, n_vars)
    tabs = [rand_fr(61 + i, 1 << n_vars) for i in range(3 if with_c else 2)]
    sq = OP.GruenSplitEq(tau)
    E_out, E_in, _ = sq.window_tables()
    got = fr_ints(R.sc_gruen_round([fr_arr(t) for t in tabs], fr_arr(E_out), fr_arr(E_in), with_c))
    assert got == OS.round_gruen_ab(tabs, sq, with_c)


def test_transcript_and_full_prove():
    a, b = R.Transcript(b"bench"), OT.Transcript(b"bench")
    vals = rand_fr(70, 5)
    for i, v in enumerate(vals):
        a.append_scalar(b"round_poly_%d" % i, fr_arr([v]))
        b.append_scalar(b"round_poly_%d" % i, v)
    assert fr_ints(a.challenge_scalar(b"round")) == [b.challenge_scalar(b"round")]
    a.append_bytes(bytes(range(200)))
    b.append_bytes(bytes(range(200)))
    assert fr_ints(a.challenge_scalar(b"x")) == [b.challenge_scalar(b"x")]
    # whole degree-2 sumcheck (config-2 shape) with the Keccak transcript
    log_n = 6
    eq = OP.eq_evals(rand_fr(71, log_n))
    m = rand_fr(72, 1 << log_n)
    evals, chal, _ = R.sc_prove("prod2", 0, [fr_arr(eq), fr_arr(m)], b"bench")
    tr = OT.Transcript(b"bench")
    tabs = [eq, m]
    for rnd in range(log_n):
        e = OS.round_prod2(tabs, 0)
        assert fr_ints(evals[rnd]) == e
        for j, v in enumerate(e[:3]):  # degree + 1 evaluations
            tr.append_scalar(b"round_poly_%d" % j, v)
        r = tr.challenge_scalar(b"round")
        assert fr_ints(chal[rnd]) == [r]
        tabs = OS.bind_tables(tabs, r, 0)
