"""GPU parity: HyperKZG.open (src/poly/commitment/mod.zig:261-324) -- SURVEY 8f row 1 -- through the C ABI vs the oracle
restatement (literal MSM commits at small sizes, the mock-SRS closed form (sum q_j tau^j) G at larger ones).
Bit-exact affine Montgomery limbs, infinity flags and final_eval."""
import numpy as np
import pytest

import zolt_b200 as Z
from oracle import bn254 as B
from oracle import poly as OP
from tests.util import P, fr_arr, fr_ints, rand_fr

pytestmark = pytest.mark.gpu


def _check(ctx, params, powers, evals, point):
    qxy, qinf, final = Z.HyperKZG.open(params, fr_arr(evals) if evals else np.empty((0, 4), np.uint64),
                                       fr_arr(point) if point else np.empty((0, 4), np.uint64))
    want_q, want_final = B.hyperkzg_open(powers, evals, point, 0)
    assert len(qxy) == len(want_q)
    got_q = B.arrays_to_points(qxy, qinf) if len(qxy) else []
    assert got_q == want_q
    if point:
        assert fr_ints(final) == [want_final]
    return got_q, fr_ints(final)[0]


@pytest.mark.parametrize("num_vars", [1, 2, 3, 5])
def test_open_vs_literal_oracle(ctx, num_vars):
    n = 1 << num_vars
    params = Z.HyperKZG.setup(ctx, n)
    powers = B.hyperkzg_setup(n)
    evals, point = rand_fr(70 + num_vars, n), rand_fr(80 + num_vars, num_vars)
    _, final = _check(ctx, params, powers, evals, point)
    # final_eval is the multilinear extension at the point, binding the TOP variable first
    cur = evals
    for r in point:
        cur = OP.bind_first(cur, r)
    assert final == cur[0]
    params.free()


@pytest.mark.parametrize("num_vars", [10, 14])
def test_open_closed_form(ctx, num_vars):
    n = 1 << num_vars
    params = Z.HyperKZG.setup(ctx, n)
    evals, point = rand_fr(90 + num_vars, n), rand_fr(91 + num_vars, num_vars)
    evals[: n // 4] = [v % 256 for v in evals[: n // 4]]  # small trace-like values
    _check(ctx, params, None, evals, point)
    # device-resident table, precomputed window-level table on the SRS: same proof
    plain = Z.HyperKZG.open(params, fr_arr(evals), fr_arr(point))
    params.precompute(0)
    d = Z.DeviceBuffer.from_host(ctx, fr_arr(evals))
    qxy, qinf, final = Z.HyperKZG.open_dev(params, d.ptr, n, fr_arr(point))
    assert (qxy == plain[0]).all() and (qinf == plain[1]).all() and (final == plain[2]).all()
    assert (d.download((n, 4)) == fr_arr(evals)).all()  # the borrowed table is not modified
    d.free()
    params.free()


def test_open_edge_cases(ctx):
    params = Z.HyperKZG.setup(ctx, 8)
    powers = B.hyperkzg_setup(8)
    # no variables: nothing to fold, no commitments
    qxy, qinf, _ = Z.HyperKZG.open(params, fr_arr([5]), np.empty((0, 4), np.uint64))
    assert len(qxy) == 0
    # table shorter than 2^num_vars: folding stops when it runs out (mod.zig:289)
    _check(ctx, params, powers, rand_fr(7, 4), rand_fr(8, 5))
    # constant table: every quotient is zero -> commitments are the identity, final_eval the constant
    q, final = _check(ctx, params, powers, [9] * 8, rand_fr(9, 3))
    assert q == [None, None, None] and final == 9
    # SRS shorter than the quotient: commit() uses the first len(powers) entries (mod.zig:241-249)
    short = Z.HyperKZG.setup(ctx, 2)
    _check(ctx, short, B.hyperkzg_setup(2), rand_fr(10, 8), rand_fr(11, 3))
    short.free()
    params.free()


def _check_batch(ctx, params, powers, polys, point, device=False):
    if device:
        bufs = [Z.DeviceBuffer.from_host(ctx, fr_arr(p)) for p in polys]
        got = Z.HyperKZG.batch_open(params, [b.ptr for b in bufs], fr_arr(point) if point else [], device_lens=[len(p) for p in polys])
        for b, p in zip(bufs, polys):
            assert (b.download((len(p), 4)) == fr_arr(p)).all()  # borrowed tables are not modified
            b.free()
    else:
        got = Z.HyperKZG.batch_open(params, [fr_arr(p) if p else np.empty((0, 4), np.uint64) for p in polys], fr_arr(point) if point else [])
    ev, qxy, qinf, final, gamma = got
    want_q, want_ev, want_final, want_gamma = B.hyperkzg_batch_open(powers, polys, point)
    assert fr_ints(gamma) == [want_gamma]
    assert fr_ints(ev) == want_ev if polys else len(ev) == 0
    assert (B.arrays_to_points(qxy, qinf) if len(qxy) else []) == want_q
    assert fr_ints(final) == [want_final]


def test_batch_open_vs_oracle(ctx):
    """HyperKZG.batchOpen (src/poly/commitment/mod.zig:607-732) incl. evaluateMultilinear as written (:788-817)."""
    params = Z.HyperKZG.setup(ctx, 64)
    powers = B.hyperkzg_setup(64)
    nv = 4
    polys = [rand_fr(500 + i, 1 << nv) for i in range(3)]
    point = rand_fr(510, nv)
    _check_batch(ctx, params, powers, polys, point)
    _check_batch(ctx, params, powers, polys, point, device=True)
    # evaluations are the multilinear extensions with point[j] <-> index bit j
    ev = Z.HyperKZG.batch_open(params, [fr_arr(p) for p in polys], fr_arr(point))[0]
    for p, e in zip(polys, fr_ints(ev)):
        cur = p
        for r in point:
            cur = OP.bind_low(cur, r)
        assert e == cur[0]
    # one polynomial; later polynomials shorter / longer than the first; an empty one
    _check_batch(ctx, params, powers, polys[:1], point)
    _check_batch(ctx, params, powers, [polys[0], polys[1][:8], polys[2] + rand_fr(511, 16), []], point)
    # no variables: no quotients, final_eval = sum gamma^i evaluation_i; and no polynomials at all
    _check_batch(ctx, params, powers, [[7], [9], [11]], [])
    _check_batch(ctx, params, powers, [], rand_fr(512, 3))
    params.free()


def test_batch_open_large_uses_the_reference_fallback(ctx):
    """Above 1024 entries (or 10 variables) evaluateMultilinear returns evals[0] in the reference: reproduced as written;
    the combined table and the proof are unaffected.  2^12 entries, closed-form commitments."""
    nv = 12
    params = Z.HyperKZG.setup(ctx, 1 << nv)
    polys = [rand_fr(520 + i, 1 << nv) for i in range(2)]
    point = rand_fr(530, nv)
    ev = Z.HyperKZG.batch_open(params, [fr_arr(p) for p in polys], fr_arr(point))[0]
    assert fr_ints(ev) == [p[0] for p in polys]
    _check_batch(ctx, params, None, polys, point)
    params.free()
