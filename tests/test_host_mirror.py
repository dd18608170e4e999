"""CPU tier: the C++ host mirror's host-only pieces (Keccak transcript, Fr glue, UniPoly) against
the oracle.  No GPU call is made here."""
import numpy as np

from oracle import bn254 as B
from oracle import poly as OP
from oracle import transcript as OT
from tests.util import P, fr_arr, fr_ints, rand_fr


def test_transcript_matches_reference_restatement(built_lib):
    from zolt_b200 import host as H

    a, b = H.Transcript(b"Jolt"), OT.Transcript(b"Jolt")
    vals = rand_fr(1, 40)
    for i, v in enumerate(vals):  # crosses the 136-byte rate several times
        a.append_scalar(b"round_poly_%d" % (i % 3), fr_arr([v]))
        b.append_scalar(b"round_poly_%d" % (i % 3), v)
        if i % 5 == 4:
            assert fr_ints(a.challenge_scalar(b"spartan_round")) == [b.challenge_scalar(b"spartan_round")]
    a.append_bytes(bytes(range(256)) * 3)
    b.append_bytes(bytes(range(256)) * 3)
    assert fr_ints(a.challenge_scalar(b"")) == [b.challenge_scalar(b"")]


def test_host_fr_glue(built_lib):
    from zolt_b200 import host as H

    n = 500
    av, bv = rand_fr(2, n), rand_fr(3, n)
    av[:4], bv[:4] = [0, 1, P - 1, P - 1], [0, P - 1, P - 1, 1]
    a, b = fr_arr(av), fr_arr(bv)
    assert fr_ints(H.fr_op(0, a, b)) == [x * y % P for x, y in zip(av, bv)]
    assert fr_ints(H.fr_op(1, a, b)) == [(x + y) % P for x, y in zip(av, bv)]
    assert fr_ints(H.fr_op(2, a, b)) == [(x - y) % P for x, y in zip(av, bv)]
    assert fr_ints(H.fr_op(6, a)) == [(-x) % P for x in av]
    assert fr_ints(H.fr_op(3, a[:50])) == [0 if x == 0 else pow(x, -1, P) for x in av[:50]]
    raw = np.random.default_rng(4).integers(0, 1 << 64, size=(200, 4), dtype=np.uint64)
    raw[0] = np.uint64(0xFFFFFFFFFFFFFFFF)
    assert fr_ints(H.fr_op(4, raw)) == [v % P for v in B.raw_array_to_ints(raw)]          # fromBytes
    assert B.raw_array_to_ints(H.fr_op(5, a)) == av                                        # toBytes (canonical)


def test_unipoly_glue(built_lib):
    from zolt_b200 import host as H

    assert fr_ints(H.evals_to_compressed(fr_arr([1, 10, 49, 142]))) == [1, 3, 4]  # src/poly/mod.zig:755-795
    e = rand_fr(5, 4)
    x = rand_fr(6, 1)
    assert fr_ints(H.evals_to_compressed(fr_arr(e))) == OP.evals_to_compressed(e)
    assert fr_ints(H.cubic_eval(fr_arr(e), fr_arr(x))) == [OP.evaluate_cubic_at_point(e, x[0])]


def test_eq_mle_and_horner_glue(built_lib):
    """EqPolynomial.mle (src/poly/mod.zig:311-321) and BatchOps.hornerEval (src/field/mod.zig:1164-1280) of the C++ mirror;
    the reference's own example: 1 + 2x + 3x^2 at x = 4 -> 57 ... and 49 for [1, 2, 3] at 4 is 1 + 8 + 48 = 57."""
    from zolt_b200 import host as H

    r, x = rand_fr(7, 9), rand_fr(8, 9)
    want = 1
    for a, b in zip(r, x):
        want = want * (a * b + (1 - a) * (1 - b)) % P
    assert fr_ints(H.eq_mle(fr_arr(r), fr_arr(x))) == [want]
    # eq on the boolean cube is the indicator
    bits = [0, 1, 1, 0, 1]
    assert fr_ints(H.eq_mle(fr_arr(bits), fr_arr(bits))) == [1]
    assert fr_ints(H.eq_mle(fr_arr(bits), fr_arr([1, 1, 1, 0, 1]))) == [0]
    coeffs, pt = rand_fr(9, 13), rand_fr(10, 1)[0]
    assert fr_ints(H.horner_eval(fr_arr(coeffs), fr_arr([pt]))) == [sum(c * pow(pt, i, P) for i, c in enumerate(coeffs)) % P]
    assert fr_ints(H.horner_eval(fr_arr([1, 2, 3]), fr_arr([4]))) == [57]
