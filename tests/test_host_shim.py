"""CPU tier: the DEVICE headers (zolt_b200/csrc/field.cuh, curve.cuh) compiled with their host
back-end -- an exact emulation of the PTX carry-flag semantics -- against the big-integer oracle.
This validates the algorithms (carry chains, even/odd accumulators, XYZZ special cases) on a box
without a GPU; the -m gpu tier checks the same functions as real PTX."""
import ctypes as C

import numpy as np
import pytest

from oracle import bn254 as B
from tests.util import rand_fr


def _run(shim, field, op, a, b):
    out = np.zeros_like(a)
    shim.shim_field_op(field, op, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_size_t(len(a)))
    return out


@pytest.mark.parametrize("field,mod", [(0, B.FR_MOD), (1, B.FQ_MOD)])
def test_field_algorithms(host_shim, field, mod):
    rng = np.random.default_rng(1)
    n = 20000
    av = [v % mod for v in B.random_fr(rng, n)]
    bv = [v % mod for v in B.random_fr(rng, n)]
    av[:6], bv[:6] = [0, 1, mod - 1, mod - 1, 0, mod - 2], [0, mod - 1, mod - 1, 1, 5, mod - 1]
    a, b = B.ints_to_raw_array(av), B.ints_to_raw_array(bv)
    rinv = pow(1 << 256, -1, mod)
    assert B.raw_array_to_ints(_run(host_shim, field, 0, a, b)) == [x * y * rinv % mod for x, y in zip(av, bv)]
    assert B.raw_array_to_ints(_run(host_shim, field, 1, a, b)) == [(x + y) % mod for x, y in zip(av, bv)]
    assert B.raw_array_to_ints(_run(host_shim, field, 2, a, b)) == [(x - y) % mod for x, y in zip(av, bv)]
    assert B.raw_array_to_ints(_run(host_shim, field, 6, a, b)) == [(-x) % mod for x in av]
    assert B.raw_array_to_ints(_run(host_shim, field, 4, a, b)) == [x * rinv % mod for x in av]
    raw = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    raw[0] = np.uint64(0xFFFFFFFFFFFFFFFF)
    assert B.raw_array_to_ints(_run(host_shim, field, 5, raw, raw)) == [(x << 256) % mod for x in B.raw_array_to_ints(raw)]
    R = (1 << 256) % mod
    got = B.raw_array_to_ints(_run(host_shim, field, 3, a[:100], b[:100]))
    assert got == [0 if x == 0 else pow(x * rinv % mod, -1, mod) * R % mod for x in av[:100]]
    # dedicated squaring (36 instead of 64 partial products): same canonical value as mul(a, a), incl. all-ones limbs
    av3 = list(av)
    av3[6:14] = [2, mod - 3, (1 << 253) + (1 << 31), 0xFFFFFFFF, (0xFFFFFFFF << 32) | 0x80000000, (1 << 254) - 1 - (1 << 254) % 1 if (1 << 254) - 1 < mod else mod - 5,
                 int("80000000" * 7, 16), int("7fffffff" + "ffffffff" * 6, 16)]
    av3 = [v % mod for v in av3]
    a3 = B.ints_to_raw_array(av3)
    assert B.raw_array_to_ints(_run(host_shim, field, 8, a3, a3)) == [x * x * rinv % mod for x in av3]
    # binary-GCD inversion (one thread normalising one MSM result): same canonical value, incl. 0 -> 0, 1, p-1, 2^k
    av2 = av[:3000]
    av2[6:12] = [2, 4, 1 << 253, (1 << 200) + 1, 3, mod - 3]
    a2 = B.ints_to_raw_array(av2)
    got = B.raw_array_to_ints(_run(host_shim, field, 7, a2, a2))
    assert got == [0 if x == 0 else pow(x * rinv % mod, -1, mod) * R % mod for x in av2]


def _fold(shim, mode, pts, neg=None, dbl=0):
    xy, _ = B.points_to_arrays(pts)
    out = np.zeros(8, dtype=np.uint64)
    oinf = C.c_uint8()
    negp = None if neg is None else np.asarray(neg, dtype=np.uint8).ctypes.data_as(C.c_void_p)
    shim.shim_g1_fold(mode, xy.ctypes.data_as(C.c_void_p), negp, C.c_size_t(len(pts)), C.c_uint32(dbl),
                      out.ctypes.data_as(C.c_void_p), C.byref(oinf))
    return B.arrays_to_points(out.reshape(1, 8), [oinf.value])[0]


def test_curve_formulas(host_shim):
    logs = rand_fr(11, 24)
    pts = [B.g1_mul(B.G1_GEN, k) for k in logs]
    total = B.g1_mul(B.G1_GEN, sum(logs))
    assert _fold(host_shim, 0, pts) == total          # mixed additions
    assert _fold(host_shim, 4, pts) == total          # warp-uniform mixed additions
    assert _fold(host_shim, 1, pts) == total          # full additions (tree)
    assert _fold(host_shim, 2, pts, dbl=5) == B.g1_mul(total, 32)
    assert _fold(host_shim, 3, pts, dbl=1000) == B.g1_mul(total, 1000)
    assert _fold(host_shim, 3, pts, dbl=0) is None
    neg = [i % 2 for i in range(24)]
    exp = B.g1_mul(B.G1_GEN, sum(-k if s else k for k, s in zip(logs, neg)))
    assert _fold(host_shim, 0, pts, neg) == exp


def test_curve_special_cases(host_shim):
    p = B.g1_mul(B.G1_GEN, 77)
    assert _fold(host_shim, 0, []) is None
    assert _fold(host_shim, 0, [p, p]) == B.g1_mul(B.G1_GEN, 154)            # P + P -> doubling (mixed)
    assert _fold(host_shim, 1, [p, p, p, p]) == B.g1_mul(B.G1_GEN, 308)      # P + P -> doubling (full)
    assert _fold(host_shim, 0, [p, p], neg=[0, 1]) is None                   # P + (-P) -> identity
    assert _fold(host_shim, 1, [p, p], neg=[0, 1]) is None
    assert _fold(host_shim, 0, [p, p, p], neg=[0, 1, 0]) == p               # identity + P
    assert _fold(host_shim, 4, [p, p]) == B.g1_mul(B.G1_GEN, 154)
    assert _fold(host_shim, 4, [p, p], neg=[0, 1]) is None
    assert _fold(host_shim, 4, [p, p, p], neg=[0, 1, 0]) == p
    assert _fold(host_shim, 4, [p, p, p, p]) == B.g1_mul(B.G1_GEN, 308)
    assert _fold(host_shim, 1, [p, p, p, p], neg=[0, 1, 1, 0]) is None      # identity + identity


@pytest.mark.parametrize("field,mod", [(0, B.FR_MOD), (1, B.FQ_MOD)])
def test_wide_accumulator(host_shim, field, mod):
    """Delayed reduction: the sum of n unreduced 512-bit products, reduced once, equals the sum of the
    individually reduced Montgomery products.  Includes the extreme operands (p-1 and lazy values up to
    4p-1) that drive the accumulator to its 2^511 normalisation threshold."""
    rng = np.random.default_rng(5)
    rinv = pow(1 << 256, -1, mod)
    for n, kind in [(1, "rand"), (7, "rand"), (300, "rand"), (300, "max"), (64, "lazy")]:
        if kind == "rand":
            av = [v % mod for v in B.random_fr(rng, n)]
            bv = [v % mod for v in B.random_fr(rng, n)]
        elif kind == "max":
            av, bv = [mod - 1] * n, [mod - 1 - i for i in range(n)]
        else:
            av, bv = [mod - 1 - i for i in range(n)], [4 * mod - 1 - i for i in range(n)]
        a, b = B.ints_to_raw_array(av), B.ints_to_raw_array(bv)
        out = np.zeros(4, dtype=np.uint64)
        host_shim.shim_wide_dot(field, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), C.c_size_t(n),
                                out.ctypes.data_as(C.c_void_p))
        want = sum(x * y for x, y in zip(av, bv)) * rinv % mod
        assert B.raw_array_to_ints(out.reshape(1, 4)) == [want], (n, kind)


@pytest.mark.parametrize("field,mod", [(0, B.FR_MOD), (1, B.FQ_MOD)])
def test_fixed_multiplier_bind(host_shim, field, mod):
    """fp_bind_fixed (the streaming sumcheck kernels' bind: table of r * 2^(32 i) * 2^64 mod p, plain 8 x 8 product, the
    addend shifted up two limbs, two Montgomery reduction rows -- the Barrett form of r2 v7 behind -DZOLT_BIND_BARRETT) returns the canonical lo + r (hi - lo), i.e. exactly what the Montgomery-product form returns, for random and
    extreme operands (0, p-1, all-ones limbs in the difference, r = 0, 1, p-1)."""
    rng = np.random.default_rng(11)
    rinv = pow(1 << 256, -1, mod)
    n = 400
    los = [v % mod for v in B.random_fr(rng, n)]
    his = [v % mod for v in B.random_fr(rng, n)]
    edge = [0, 1, mod - 1, mod - 2, (1 << 253), (1 << 224) - 1, mod >> 1]
    for i, a in enumerate(edge):
        for j, b in enumerate(edge):
            los[(7 * i + j) % n], his[(7 * i + j) % n] = a, b
    lo, hi = B.ints_to_raw_array(los), B.ints_to_raw_array(his)
    for r in [0, 1, mod - 1, (1 << 256) % mod, (1 << 253) + 5, 0xFFFFFFFF, ((1 << 125) - 1) << 128] + [v % mod for v in B.random_fr(rng, 6)]:
        rr = B.ints_to_raw_array([r % mod])
        out = np.zeros((n, 4), dtype=np.uint64)
        host_shim.shim_bind_fixed(field, rr.ctypes.data_as(C.c_void_p), lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p),
                                  C.c_size_t(n), out.ctypes.data_as(C.c_void_p))
        r_std = (r % mod) * rinv % mod
        want = [(a + r_std * (b - a)) % mod for a, b in zip(los, his)]
        assert B.raw_array_to_ints(out) == want, hex(r)
        # the FP64-pipe form (16-bit x 32-bit limb products summed exactly in doubles, one-word Barrett): same values
        out64 = np.zeros((n, 4), dtype=np.uint64)
        host_shim.shim_bind_fixed_f64(field, rr.ctypes.data_as(C.c_void_p), lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p),
                                      C.c_size_t(n), out64.ctypes.data_as(C.c_void_p))
        assert B.raw_array_to_ints(out64) == want, ("f64", hex(r))


@pytest.mark.parametrize("field,mod", [(0, B.FR_MOD), (1, B.FQ_MOD)])
def test_two_products_one_reduction(host_shim, field, mod):
    """fp_mul2_sub (y3 = r (q - x3) - y1 ppp of the MSM's mixed addition with ONE Montgomery reduction): equal to the
    difference of the two reduced products for random operands and for the extremes (0, 1, p-1 in every position, which
    drive the shared accumulator to its 3p bound and the subtrahend's complement to p itself)."""
    rng = np.random.default_rng(21)
    n = 4000
    vals = [[v % mod for v in B.random_fr(rng, n)] for _ in range(4)]
    ext = [0, 1, mod - 1, mod - 2, (1 << 253), mod >> 1]
    k = 0
    for a in ext:
        for b in ext:
            for c in ext:
                for d in (0, mod - 1, 1):
                    vals[0][k], vals[1][k], vals[2][k], vals[3][k] = a, b, c, d
                    k += 1
    arrs = [B.ints_to_raw_array(v) for v in vals]
    out = np.zeros((n, 4), dtype=np.uint64)
    host_shim.shim_mul2_sub(field, *[x.ctypes.data_as(C.c_void_p) for x in arrs], C.c_size_t(n), out.ctypes.data_as(C.c_void_p))
    rinv = pow(1 << 256, -1, mod)
    want = [(a * b - c * d) * rinv % mod for a, b, c, d in zip(*vals)]
    assert B.raw_array_to_ints(out) == want


# ---- property tests (hypothesis): limb-pattern-biased operands for the arithmetic added in r1 -----------------------
from hypothesis import given, settings
from hypothesis import strategies as st


def _limb_biased(mod):
    """Field elements whose 32-bit limbs are drawn from {0, 1, 2^31, 2^32-1, random}: carries and all-ones runs."""
    limb = st.one_of(st.sampled_from([0, 1, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF, 0xFFFFFFFE]), st.integers(0, 0xFFFFFFFF))
    return st.lists(limb, min_size=8, max_size=8).map(lambda ls: sum(v << (32 * i) for i, v in enumerate(ls)) % mod)


@settings(max_examples=300, deadline=None)
@given(data=st.data())
def test_property_sqr_mul2_bind(host_shim, data):
    field = data.draw(st.sampled_from([0, 1]))
    mod = B.FR_MOD if field == 0 else B.FQ_MOD
    a, b, c, d = (data.draw(_limb_biased(mod)) for _ in range(4))
    rinv = pow(1 << 256, -1, mod)
    arr = lambda v: B.ints_to_raw_array([v])  # noqa: E731
    out = np.zeros((1, 4), dtype=np.uint64)
    p = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
    A, Bv, Cv, D = arr(a), arr(b), arr(c), arr(d)
    # squaring
    host_shim.shim_field_op(field, 8, p(A), p(A), p(out), C.c_size_t(1))
    assert B.raw_array_to_ints(out) == [a * a * rinv % mod]
    # two products, one reduction
    host_shim.shim_mul2_sub(field, p(A), p(Bv), p(Cv), p(D), C.c_size_t(1), p(out))
    assert B.raw_array_to_ints(out) == [(a * b - c * d) * rinv % mod]
    # bind through the multiplier table: lo + r (hi - lo), r in Montgomery form
    host_shim.shim_bind_fixed(field, p(A), p(Bv), p(Cv), C.c_size_t(1), p(out))
    assert B.raw_array_to_ints(out) == [(b + a * rinv % mod * (c - b)) % mod]
    host_shim.shim_bind_fixed_f64(field, p(A), p(Bv), p(Cv), C.c_size_t(1), p(out))
    assert B.raw_array_to_ints(out) == [(b + a * rinv % mod * (c - b)) % mod]
    # binary-GCD inverse
    host_shim.shim_field_op(field, 7, p(A), p(A), p(out), C.c_size_t(1))
    R = (1 << 256) % mod
    assert B.raw_array_to_ints(out) == [0 if a == 0 else pow(a * rinv % mod, -1, mod) * R % mod]
