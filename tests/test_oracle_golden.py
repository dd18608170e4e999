"""CPU tier: pin the oracle against the reference's own golden vectors.

  * logs/zolt.log claim-update KATs (committed as tests/golden/zolt_log_kats.json by make_golden.py)
  * the exact values asserted by the reference's unit tests (file:line cited per test)
  * the Montgomery constants of src/field/mod.zig:16-75
"""
import json
from pathlib import Path

from oracle import bn254 as B
from oracle import poly as OP
from oracle import sumcheck as OS
from oracle import transcript as OT

P = B.FR_MOD
GOLD = Path(__file__).resolve().parent / "golden"


def test_log_claim_update_kats():
    kats = json.loads((GOLD / "zolt_log_kats.json").read_text())["kats"]
    assert len(kats) >= 20
    le = lambda b: int.from_bytes(bytes(b), "little")  # noqa: E731
    for k in kats:
        got = OP.compressed_claim_update(le(k["current_claim"]), le(k["c0"]), le(k["c2"]), le(k["c3"]), le(k["challenge"]))
        assert got == le(k["next_claim"]), (k["stage"], k["round"])
        assert B.fr_to_bytes(got) == bytes(k["next_claim"])  # toBytes = canonical LE (field/mod.zig:685-695)


def test_log_stage3_round_kats():
    """The eight STAGE3_ROUND_k tuples generateStage3Proof printed in the captured run (stage3_prover.zig:430-560): the batched
    cubic's p(0) is c0, p(0) + p(1) is the running claim, and the claim after the challenge is what the evaluation-form path
    of the Stage 3 mirror (evaluate_cubic_at_point over [p(0..3)]) and the compressed-coefficient path both return."""
    from oracle import stage3 as S3

    kats = [k for k in json.loads((GOLD / "zolt_log_kats.json").read_text())["kats"] if k["stage"] == "STAGE3"]
    assert len(kats) == 8  # trace_len 256 -> 8 rounds
    le = lambda b: int.from_bytes(bytes(b), "little")  # noqa: E731
    prev_next = None
    for k in sorted(kats, key=lambda x: x["round"]):
        claim, c0, c2, c3, r = (le(k[x]) for x in ("current_claim", "c0", "c2", "c3", "challenge"))
        if "p0" in k:  # printed for the first rounds only
            assert le(k["p0"]) == c0 and (le(k["p0"]) + le(k["p1"])) % P == claim
        if prev_next is not None:
            assert claim == prev_next  # the rounds chain
        c1 = (claim - 2 * c0 - c2 - c3) % P
        evals = [(c0 + c1 * x + c2 * x * x + c3 * x ** 3) % P for x in range(4)]
        if "p1" in k:
            assert evals[1] == le(k["p1"])
        assert S3._cubic_at(evals, r) == le(k["next_claim"])
        # the Shift instance's own p(0), p(1) of the same round are field elements below r
        assert all(le(k[x]) < P for x in ("shift_p0", "shift_p1") if x in k)
        prev_next = le(k["next_claim"])
    # the preamble of the same run: Shift's first round sums to its input claim (the identity oracle/stage3.py's phase 1
    # relies on), and the Shift gamma powers are powers of one gamma (what the Stage 3 tests feed the provers)
    pre = json.loads((GOLD / "zolt_log_kats.json").read_text())["stage3_pre"]
    r0 = min(kats, key=lambda x: x["round"])
    assert (le(r0["shift_p0"]) + le(r0["shift_p1"])) % P == le(pre["input_claim_0"])
    assert le(pre["gamma_power_0"]) == 1 and le(pre["gamma_power_4"]) == pow(le(pre["gamma_power_1"]), 4, P)


def test_montgomery_constants_match_reference_source():
    # src/field/mod.zig:24-41 (Fr) and :59-75 (Fq)
    assert B.int_to_limbs(B.FR_R) == [0xAC96341C4FFFFFFB, 0x36FC76959F60CD29, 0x666EA36F7879462E, 0x0E0A77C19A07DF2F]
    assert B.int_to_limbs(B.FR_R2) == [0x1BB8E645AE216DA7, 0x53FE3AB1E35C59E3, 0x8C49833D53BB8085, 0x0216D0B17F4E44A5]
    assert B.FR_INV64 == 0xC2E1F593EFFFFFFF
    assert B.int_to_limbs(B.FQ_R) == [0xD35D438DC58F0D9D, 0x0A78EB28F5C70B3D, 0x666EA36F7879462C, 0x0E0A77C19A07DF2F]
    assert B.int_to_limbs(B.FQ_R2) == [0xF32CFC5B538AFA89, 0xB5E71911D44501FB, 0x47AB1EFF0A417FF6, 0x06D89F71CAB8351F]
    assert B.FQ_INV64 == 0x87D20782E4866389


def test_bind_low_unit_vectors():
    # src/poly/mod.zig:816-888
    assert OP.bind_low([1, 2, 3, 4], 3) == [4, 6]
    assert OP.bind_low([10, 20, 30, 40, 50, 60, 70, 80], 5) == [60, 80, 100, 120]


def test_half_split_sumcheck_unit_vectors():
    # src/subprotocols/mod.zig:379-436: g(0)=3, g(1)=7, after r=2 -> [5,6]; :441-461 total 36
    assert OS.round_sum([[1, 2, 3, 4]], OS.HALF) == [3, 7]
    assert OP.bind_first([1, 2, 3, 4], 2) == [5, 6]
    assert sum(range(1, 9)) == 36 and sum(OS.round_sum([list(range(1, 9))], OS.HALF)) == 36


def test_interpolation_unit_vectors():
    # src/poly/mod.zig:755-795: 1 + 2X + 3X^2 + 4X^3 from [1, 10, 49, 142]; compressed [1, 3, 4]
    assert OP.interpolate_degree3([1, 10, 49, 142]) == [1, 2, 3, 4]
    assert OP.evals_to_compressed([1, 10, 49, 142]) == [1, 3, 4]
    # cubic evaluation (src/zkvm/batched_sumcheck.zig:475-494)
    for x in (0, 1, 2, 3, 5, 12345):
        assert OP.evaluate_cubic_at_point([1, 10, 49, 142], x) == (1 + 2 * x + 3 * x * x + 4 * x**3) % P


def test_eq_partition_of_unity_and_big_endian():
    # src/poly/mod.zig:689-753
    r = [3, 5, 7]
    ev = OP.eq_evals(r)
    assert sum(ev) % P == 1
    for idx in range(8):
        bits = [(idx >> (2 - j)) & 1 for j in range(3)]  # r[0] <-> MSB
        assert ev[idx] == OP.eq_mle(r, bits)


def test_split_eq_tables_unit_vectors():
    # src/poly/split_eq.zig:575-676: tau=[2,3,5] -> m=1, E_out over tau[0], E_in over tau[1]; eq(3,5)=23 (:554-573)
    sq = OP.GruenSplitEq([2, 3, 5])
    assert sq.E_out_vec == [[1], [(1 - 2) % P, 2]]
    assert sq.E_in_vec == [[1], [(1 - 3) % P, 3]]
    sq4 = OP.GruenSplitEq([3, 5, 7, 11])
    assert len(sq4.E_out_vec) == 3 and len(sq4.E_in_vec) == 2
    one = OP.GruenSplitEq([3])
    one.bind(5)
    assert one.current_scalar == (3 * 5 + (1 - 3) * (1 - 5)) % P == 23


def test_gruen_cubic_is_consistent_with_dense_eq():
    # s(X) from computeCubicRoundPoly must equal the dense sum eq(tau,x) l(x) r(x) round polynomial
    import numpy as np

    rng = np.random.default_rng(5)
    n_vars = 5
    tau = B.random_fr(rng, n_vars)
    left, right = B.random_fr(rng, 1 << n_vars), B.random_fr(rng, 1 << n_vars)
    eq = OP.eq_evals(tau)
    sq = OP.GruenSplitEq(tau)
    claim = sum(e * a * b for e, a, b in zip(eq, left, right)) % P
    tabs = [left, right]
    dense = [eq, left, right]
    for _ in range(n_vars):
        t0, tinf = OS.round_gruen_ab(tabs, sq)
        s = sq.cubic_round_poly(t0, tinf, claim)
        assert s == OS.round_prod3(dense, OS.ADJACENT)
        assert (s[0] + s[1]) % P == claim
        r = B.random_fr(rng, 1)[0]
        claim = OP.evaluate_cubic_at_point(s, r)
        tabs = OS.bind_tables(tabs, r, OS.ADJACENT)
        dense = OS.bind_tables(dense, r, OS.ADJACENT)
        sq.bind(r)


def test_field_small_integer_identities():
    # src/field/mod.zig:1101-1140, :1452-1494: 3*7=21, 2^3=8, inner product 70, Horner 49
    m = lambda v: B.fr_to_mont(v)  # noqa: E731
    assert B.fr_from_mont(B.mont_mul(m(3), m(7), P)) == 21
    assert B.fr_from_mont(B.mont_mul(B.mont_mul(m(2), m(2), P), m(2), P)) == 8
    assert sum(a * b for a, b in zip([1, 2, 3, 4], [5, 6, 7, 8])) % P == 70
    assert OP.unipoly_evaluate([1, 2, 3, 4], 2) == 49


def test_window_digits_and_sizes():
    # src/msm/mod.zig:827-851
    assert [B.optimal_window_size(n) for n in (1, 7, 8, 31, 32, 127, 128, 511, 512, 2047, 2048, 8191, 8192, 32767, 32768, 1 << 24)] == [
        1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8]
    s = 0b11001010
    assert [B.get_window(s, w, 2) for w in range(4)] == [0b10, 0b10, 0b00, 0b11]
    assert B.get_window(s, 0, 4) == 0b1010 and B.get_window(s, 1, 4) == 0b1100


def test_msm_literal_pippenger_equals_closed_form():
    # bench.zig workloads (src/bench.zig:243-287) -- scalars 7i+13, bases (i+1)G
    for n in (16, 64, 256):
        bases, acc = [], None
        for i in range(n):
            acc = B.g1_add(acc, B.G1_GEN)
            bases.append(acc)
        scalars = [7 * i + 13 for i in range(n)]
        assert B.msm(bases, scalars) == B.msm_closed_form(range(1, n + 1), scalars)
    assert B.msm([], []) is None


def test_commit_kat():
    # SURVEY 8c.3 derived KAT for the mock SRS (tau = 0x12345678, commitment/mod.zig:189)
    pt = B.hyperkzg_commit(B.hyperkzg_setup(4), [1, 2, 3, 4])
    assert pt == (
        0x07AEB450B7B66FDF44A37824224849B29C1ED87D705324299A17C8FB790F75CE,
        0x166808E4F5E0101BE322704903A3156ABBE3AD2C7A8992563F704CE020B05D14,
    )
    assert B.int_to_limbs(B.fq_to_mont(pt[0])) == [0x34686E9C2DC1A390, 0x6E093777DB13CF10, 0xAE87848108CE376F, 0x23A755E5AA7A0FD3]
    assert B.g1_is_on_curve(pt) and B.hyperkzg_setup(2)[1] == B.g1_mul(B.G1_GEN, B.HYPERKZG_TAU)


def test_transcript_is_deterministic_and_sensitive():
    # src/integration_tests.zig:80-118 (determinism); byte contract of src/transcripts/mod.zig:87-131
    a, b = OT.Transcript(b"Jolt"), OT.Transcript(b"Jolt")
    a.append_scalar(b"x", 5)
    b.append_scalar(b"x", 5)
    ca, cb = a.challenge_scalar(b"c"), b.challenge_scalar(b"c")
    assert ca == cb and 0 <= ca < P
    c = OT.Transcript(b"Jolt")
    c.append_scalar(b"x", 6)
    assert c.challenge_scalar(b"c") != ca
    # absorbing exactly the rate triggers one permutation and resets the position
    t = OT.Transcript(b"")
    t.append_bytes(bytes(136))
    assert t.position == 0 and any(t.state)
    # keccak-f[1600] on the zero state: first lane of the well-known test vector
    st = bytearray(200)
    OT.keccak_f(st)
    assert int.from_bytes(st[:8], "little") == 0xF1258F7940E1DDE7


def test_hyperkzg_open_oracle_consistency():
    """HyperKZG.open restatement (src/poly/commitment/mod.zig:261-324): the final evaluation is the multilinear
    extension at the point (top variable first) and the literal MSM commits agree with the mock-SRS closed form
    (sum_j q_j tau^j) G; edge cases of :270-276 (no variables) and :289 (table runs out)."""
    import numpy as np

    rng = np.random.default_rng(3)
    evals, point = B.random_fr(rng, 8), B.random_fr(rng, 3)
    powers = B.hyperkzg_setup(8)
    q_lit, fin = B.hyperkzg_open(powers, evals, point, 0)
    q_cf, fin2 = B.hyperkzg_open(None, evals, point, 0)
    assert q_lit == q_cf and fin == fin2 and len(q_lit) == 3
    cur = evals
    for r in point:
        cur = OP.bind_first(cur, r)
    assert fin == cur[0]
    assert B.hyperkzg_open(powers, evals, [], 42) == ([], 42)
    assert len(B.hyperkzg_open(powers, evals[:2], point, 0)[0]) == 1


def test_batched_sumcheck_restatement_final_claim():
    """BatchedSumcheckProver restatement (src/zkvm/batched_sumcheck.zig:158-257, :330-420): with instances of equal
    length the final claim is the batching-coefficient combination of the instances' combiners at the bound point; the
    claim chain uses the compressed update pinned by the logs/zolt.log KATs above."""
    import numpy as np

    rng = np.random.default_rng(1)
    inst = [("prod3", [B.random_fr(rng, 16) for _ in range(3)]), ("sum", [B.random_fr(rng, 16)])]
    coeffs, ch = B.random_fr(rng, 2), B.random_fr(rng, 4)
    out, final, claims = OS.run_batched(inst, coeffs, ch)
    _, b0 = OS.run_dense("prod3", inst[0][1], 0, ch)
    _, b1 = OS.run_dense("sum", inst[1][1], 0, ch)
    f0, f1 = [t[0] for t in b0[-1]], b1[-1][0][0]
    assert final == (coeffs[0] * f0[0] * f0[1] * f0[2] + coeffs[1] * f1) % P
    assert len(out) == 4 and all(len(c) == 3 for c in out)


def test_raf_restatement_equals_dense_product():
    """RafEvaluationProver as written (implicit unmap, src/zkvm/ram/raf_checking.zig:335-409) == the dense degree-2
    product over (ra, unmap table): the identity the device path relies on (tests/test_gpu_sumcheck.py)."""
    import numpy as np

    rng = np.random.default_rng(2)
    k, start = 16, 0x8000_0000
    ra, ch = B.random_fr(rng, k), B.random_fr(rng, 4)
    claim = sum(a * (start + 8 * i) for i, a in enumerate(ra)) % P
    ev, tabs, fin = OS.run_raf(ra, start, ch, claim)
    ev2, b2 = OS.run_dense("prod2", [ra, [start + 8 * i for i in range(k)]], 0, ch)
    assert ev == [list(e) for e in ev2]
    assert [t for t in tabs] == [b[0] for b in b2]
    assert fin == b2[-1][0][0] * b2[-1][1][0] % P


def _load_vectors():
    return json.loads((GOLD / "oracle_vectors.json").read_text())


def _ints(xs):
    return [int(x, 16) for x in xs]


def _pt(p):
    return None if p is None else (int(p[0], 16), int(p[1], 16))


def test_oracle_reproduces_committed_vectors():
    """tests/golden/oracle_vectors.json (written by make_oracle_vectors.py): the Python oracle AND the C++ restatement
    (oracle/cpu_ref.cpp) still produce the committed outputs -- a drift in either is caught here, on the CPU tier."""
    import numpy as np

    from oracle import cpu_ref as R

    v = _load_vectors()
    logs, scalars = _ints(v["msm"]["base_logs"]), _ints(v["msm"]["scalars"])
    bases = [B.g1_mul(B.G1_GEN, k) for k in logs]
    assert B.msm(bases, scalars) == _pt(v["msm"]["result"]) == B.msm_closed_form(logs, scalars)
    xy, inf = B.points_to_arrays(bases)
    cxy, cinf = R.msm(xy, inf, B.ints_to_mont_array(scalars, P), threads=1)
    assert B.arrays_to_points(cxy.reshape(1, 8), [cinf])[0] == _pt(v["msm"]["result"])
    hk = v["hyperkzg"]
    powers = B.hyperkzg_setup(8)
    assert B.hyperkzg_commit(powers, _ints(hk["evals"])) == _pt(hk["commitment"])
    q, fin = B.hyperkzg_open(powers, _ints(hk["evals"]), _ints(hk["point"]), 0)
    assert q == [_pt(p) for p in hk["quotients"]] and fin == int(hk["final_eval"], 16)
    for kind, s in v["sumcheck"].items():
        tabs = [_ints(t) for t in s["tables"]]
        ev, bound = OS.run_dense(kind, tabs, s["pairing"], _ints(s["challenges"]))
        assert [list(e) for e in ev] == [_ints(e) for e in s["round_evals"]], kind
        assert [t[0] for t in bound[-1]] == _ints(s["final_values"]), kind
    g = v["gruen_ab"]
    ev, _ = OS.run_gruen([_ints(t) for t in g["tables"]], _ints(g["tau"]), _ints(g["challenges"]), False)
    assert [list(e) for e in ev] == [_ints(e) for e in g["round_sums"]]
