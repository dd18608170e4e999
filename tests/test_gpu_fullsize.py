"""GPU parity at BASELINE.json's full sizes (configs 3 and 4: 2^24 entries / points).

The Python big-integer oracle cannot finish 2^24 in seconds, so these cases use
  * config 4 (HyperKZG commit, MSM 2^24): the closed form commit(e) == (sum_i e_i tau^i) * G for the mock SRS
    tau = 0x12345678 (src/poly/commitment/mod.zig:174-213): one Horner pass over Python integers + ONE oracle
    scalar multiplication, compared bit for bit with the affine point the device returns, for the plain path
    (16 signed windows) and the precomputed window-level table (what bench.py times);
  * config 3 (degree-3 Spartan sumcheck 2^24): the C++ restatement of the reference's round loop
    (oracle/cpu_ref.cpp, ~8 s single-threaded at this size) -- every round polynomial, every challenge and the
    final bound values bit for bit -- plus the verifier's own chain p_j(0) + p_j(1) = claim_j,
    claim_{j+1} = p_j(r_j), claim_final = eq * (a * b - c) at the bound point.
"""
import numpy as np
import pytest

import zolt_b200 as Z
from oracle import bn254 as B
from oracle import cpu_ref as R
from oracle import poly as OP
from tests.util import P, fr_ints
from zolt_b200 import host as H

pytestmark = pytest.mark.gpu

LOG_N = 24


def _rand_mont(seed, n):
    """n canonical Montgomery-form elements, uniform below 2^252 < r (synthetic data, like bench.py)."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


def _limbs_to_ints(a):
    raw = np.ascontiguousarray(a).tobytes()
    return [int.from_bytes(raw[32 * i:32 * i + 32], "little") for i in range(a.shape[0])]


def test_config4_hyperkzg_commit_2_24_closed_form(ctx):
    n = 1 << LOG_N
    tau = 0x12345678
    srs = Z.HyperKZG.setup(ctx, n)  # powers tau^i * G generated on the device
    try:
        assert len(srs) == n
        evals = _rand_mont(0x5A170004, n)
        # "real trace" flavour (src/zkvm/mod.zig:1531-1600): many zeros and small values in the first quarter
        small = np.random.default_rng(5).integers(0, 256, size=n // 4, dtype=np.uint64)
        small[::3] = 0
        byte_vals = B.ints_to_mont_array(list(range(256)), P)  # Montgomery forms of 0..255
        evals[: n // 4] = byte_vals[small]
        # closed form: Montgomery limbs hold e_i * R, so Horner over them gives R * sum e_i tau^i
        acc = 0
        for e in reversed(_limbs_to_ints(evals)):
            acc = (acc * tau + e) % P
        s = acc * pow(1 << 256, -1, P) % P
        want = B.g1_mul(B.G1_GEN, s)
        xy, inf = Z.HyperKZG.commit(srs, evals)  # plain bases: 16 signed windows
        got = B.arrays_to_points(xy.reshape(1, 8), [inf])[0]
        assert got == want, "2^24 commit (plain bases) differs from (sum e_i tau^i) G"
        srs.precompute(0)  # window-level table, the path bench.py times
        xy2, inf2 = Z.HyperKZG.commit(srs, evals)
        assert (xy2 == xy).all() and inf2 == inf, "2^24 commit (window-level table) differs from the plain path"
    finally:
        srs.free()


def test_config3_spartan_sumcheck_2_24_vs_cpu_restatement(ctx):
    n = 1 << LOG_N
    eq, az, bz = (_rand_mont(0x5A170030 + t, n) for t in range(3))
    # Cz = Az o Bz on ~90 % of the rows (a mostly satisfied R1CS), random elsewhere
    cz = np.empty_like(az)
    step = 1 << 20
    for lo in range(0, n, step):
        cz[lo:lo + step] = ctx.field_op(Z.FIELD_FR, Z.OP_MUL, az[lo:lo + step], bz[lo:lo + step])
    bad = np.random.default_rng(0x5A170033).random(n) < 0.1
    cz[bad] = _rand_mont(0x5A170034, int(bad.sum()))
    tabs = [eq, az, bz, cz]
    evals, chal, final = H.sumcheck_prove(ctx, Z.SC_EQ_AB_MINUS_C, Z.PAIR_ADJACENT, tabs)
    r_evals, r_chal, r_final = R.sc_prove("eq_ab_minus_c", 0, tabs)
    assert (chal == r_chal).all(), "challenges differ"
    assert (evals == r_evals).all(), "round polynomials differ"
    for t in range(4):
        assert (final[t] == r_final[t][0]).all(), f"final value of table {t} differs"
    # compressed form [c0, c2, c3] the Jolt path absorbs (src/poly/mod.zig:682-685), first and last round
    for j in (0, LOG_N - 1):
        e = fr_ints(evals[j])
        assert fr_ints(H.evals_to_compressed(evals[j])) == OP.evals_to_compressed(e)
    # the verifier's chain
    e0 = fr_ints(evals[0])
    claim = (e0[0] + e0[1]) % P
    for j in range(LOG_N):
        e = fr_ints(evals[j])
        assert (e[0] + e[1]) % P == claim, f"round {j}: p(0) + p(1) != claim"
        claim = OP.evaluate_cubic_at_point(e, fr_ints(chal[j])[0])
    f = [fr_ints(final[t])[0] for t in range(4)]
    assert claim == f[0] * (f[1] * f[2] - f[3]) % P


def test_config3_bound_tables_at_2_24(ctx):
    """BASELINE.md section 4's table checks at full size: after binds 0, 1, 2 a checksum of every bound table (its sum and its
    inner product with a fixed random vector, both mod r, device reductions against the threaded C++ restatement), and the full
    tables of the last 8 rounds element by element."""
    n = 1 << LOG_N
    tabs = [_rand_mont(0x5A170040 + t, n) for t in range(4)]
    chal = _rand_mont(0x5A170048, LOG_N)
    inst = Z.SumcheckInstance(ctx, Z.SC_EQ_AB_MINUS_C, Z.PAIR_ADJACENT, tabs)
    weights = _rand_mont(0x5A170049, n // 2)
    d_w = Z.DeviceBuffer.from_host(ctx, weights)
    one = B.ints_to_mont_array([1], P)
    ones = np.repeat(one, n // 2, axis=0)
    ref, ref_round = list(tabs), 0

    def advance(to):
        nonlocal ref, ref_round
        while ref_round < to:
            ref = [R.bind_low(t, chal[ref_round]) for t in ref]
            ref_round += 1

    try:
        for rnd in range(LOG_N):
            inst.compute_round_poly()
            inst.bind_challenge(chal[rnd:rnd + 1])
            cur = n >> (rnd + 1)
            if rnd < 3:
                advance(rnd + 1)
                for t in range(4):
                    ptr = inst.table_dev(t)
                    got_sum = ctx.poly_sum(ptr, cur)
                    got_dot = ctx.fr_inner_product_dev(ptr, d_w.ptr, cur)
                    assert (got_sum == R.fr_dot(ref[t], ones[:cur], threads=8)).all(), f"round {rnd} table {t}: sum"
                    assert (got_dot == R.fr_dot(ref[t], weights[:cur], threads=8)).all(), f"round {rnd} table {t}: checksum"
            elif cur <= 256:
                advance(rnd + 1)
                for t in range(4):
                    assert (inst.table(t) == ref[t]).all(), f"round {rnd} table {t}"
    finally:
        d_w.free()
        inst.free()
