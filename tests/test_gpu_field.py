"""GPU parity: Fr / Fq element-wise ops through zolt_field_batch_op vs big-integer math.
Bit-exact (integer work).  Reference: src/field/mod.zig:735-816, :944-983, :617-652."""
import numpy as np
import pytest

import zolt_b200 as Z
from oracle import bn254 as B

pytestmark = pytest.mark.gpu


def _operands(mod, n, seed):
    rng = np.random.default_rng(seed)
    a = [v % mod for v in B.random_fr(rng, n)]
    b = [v % mod for v in B.random_fr(rng, n)]
    edge = [0, 1, 2, mod - 1, mod - 2, (1 << 253), mod >> 1, B.FR_R % mod]
    for i, x in enumerate(edge):
        for j, y in enumerate(edge):
            k = i * len(edge) + j
            if k < n:
                a[k], b[k] = x, y
    return a, b


@pytest.mark.parametrize("field,mod", [(Z.FIELD_FR, B.FR_MOD), (Z.FIELD_FQ, B.FQ_MOD)])
def test_field_ops_bit_exact(ctx, field, mod):
    n = 1 << 14
    av, bv = _operands(mod, n, 7 + field)
    a, b = B.ints_to_raw_array(av), B.ints_to_raw_array(bv)
    rinv = pow(1 << 256, -1, mod)
    got = B.raw_array_to_ints(ctx.field_op(field, Z.OP_MUL, a, b))
    assert got == [x * y * rinv % mod for x, y in zip(av, bv)]
    got = B.raw_array_to_ints(ctx.field_op(field, Z.OP_ADD, a, b))
    assert got == [(x + y) % mod for x, y in zip(av, bv)]
    got = B.raw_array_to_ints(ctx.field_op(field, Z.OP_SUB, a, b))
    assert got == [(x - y) % mod for x, y in zip(av, bv)]
    got = B.raw_array_to_ints(ctx.field_op(field, Z.OP_NEG, a))
    assert got == [(-x) % mod for x in av]
    got = B.raw_array_to_ints(ctx.field_op(field, Z.OP_FROM_MONT, a))
    assert got == [x * rinv % mod for x in av]


@pytest.mark.parametrize("field,mod", [(Z.FIELD_FR, B.FR_MOD), (Z.FIELD_FQ, B.FQ_MOD)])
def test_to_mont_accepts_any_256_bit_value(ctx, field, mod):
    # fromBytes semantics (src/field/mod.zig:625-639): values >= p are reduced
    rng = np.random.default_rng(11)
    raw = rng.integers(0, 1 << 64, size=(4096, 4), dtype=np.uint64)
    raw[0] = np.uint64(0xFFFFFFFFFFFFFFFF)
    raw[1] = 0
    vals = B.raw_array_to_ints(raw)
    got = B.raw_array_to_ints(ctx.field_op(field, Z.OP_TO_MONT, raw))
    assert got == [(v << 256) % mod for v in vals]


@pytest.mark.parametrize("field,mod", [(Z.FIELD_FR, B.FR_MOD), (Z.FIELD_FQ, B.FQ_MOD)])
def test_inverse(ctx, field, mod):
    n = 256
    av, _ = _operands(mod, n, 3)
    R = (1 << 256) % mod
    rinv = pow(R, -1, mod)
    got = B.raw_array_to_ints(ctx.field_op(field, Z.OP_INV, B.ints_to_raw_array(av)))
    exp = [0 if x == 0 else pow(x * rinv % mod, -1, mod) * R % mod for x in av]
    assert got == exp


def test_reference_small_integer_identities(ctx):
    # src/field/mod.zig:1101-1140: 3*7 = 21, 2^3 = 8 in Montgomery form
    three, seven, two = (B.ints_to_mont_array([v], B.FR_MOD) for v in (3, 7, 2))
    assert B.mont_array_to_ints(ctx.field_op(Z.FIELD_FR, Z.OP_MUL, three, seven), B.FR_MOD) == [21]
    four = ctx.field_op(Z.FIELD_FR, Z.OP_MUL, two, two)
    assert B.mont_array_to_ints(ctx.field_op(Z.FIELD_FR, Z.OP_MUL, four, two), B.FR_MOD) == [8]


@pytest.mark.parametrize("n", [1, 7, 1000, 70001])
def test_inner_product_dev(ctx, n):
    """BatchOps.innerProduct (src/field/mod.zig:1164-1280) over device-resident vectors."""
    from tests.util import fr_arr, fr_ints, rand_fr

    a, b = rand_fr(9100 + n, n), rand_fr(9200 + n, n)
    a[0], b[0] = B.FR_MOD - 1, B.FR_MOD - 1
    da, db = Z.DeviceBuffer.from_host(ctx, fr_arr(a)), Z.DeviceBuffer.from_host(ctx, fr_arr(b))
    assert fr_ints(ctx.fr_inner_product_dev(da.ptr, db.ptr, n)) == [sum(x * y for x, y in zip(a, b)) % B.FR_MOD]
    da.free()
    db.free()
