"""The Stage 3 oracle (oracle/stage3.py) against the reference's own unit expectations for EqPlusOnePolynomial
(src/poly/mod.zig:890-960) and against the identities generateStage3Proof's debug lines assert on every run
(stage3_prover.zig: p(0) + p(1) = claim per prover and round, grand_sum(P*Q) = direct_sum :1131-1170, the final-claim
relations :620-700).  Also pins the closed form the device kernel uses for the eq+1 table to the literal mle."""
import random

import pytest

from oracle import stage3 as S
from oracle.bn254 import FR_MOD as P
from oracle.poly import eq_mle, evaluate_cubic_at_point


def test_eq_plus_one_reference_unit_expectations():
    # "EqPlusOnePolynomial basic": 2-bit successor relation
    for x in range(4):
        for y in range(4):
            xb, yb = [(x >> 1) & 1, x & 1], [(y >> 1) & 1, y & 1]
            assert S.eq_plus_one_mle(xb, yb) == (1 if (y == x + 1 and x < 3) else 0)


@pytest.mark.parametrize("n", [1, 2, 3, 5])
def test_eq_plus_one_table_closed_form(n):
    """entry j of the table = (1 - r'_k) prod_{i<k} r'_i prod_{i>k} (j_i ? r'_i : 1 - r'_i), k = trailing zeros of j,
    r'_i = r[n-1-i]; entry 0 = 0 -- what eq_plus_one_kernel evaluates."""
    rng = random.Random(n)
    r = [rng.randrange(P) for _ in range(n)]
    table = S.eq_plus_one_evals(r)
    assert table[0] == 0
    for j in range(1, 1 << n):
        k = (j & -j).bit_length() - 1
        acc = (1 - r[n - 1 - k]) % P
        for i in range(k):
            acc = acc * r[n - 1 - i] % P
        for i in range(k + 1, n):
            x = r[n - 1 - i]
            acc = acc * (x if (j >> i) & 1 else (1 - x) % P) % P
        assert acc == table[j]
    # a boolean point: the table is the indicator of r + 1
    for x in range(1 << n):
        t = S.eq_plus_one_evals(S._bits_be(x, n))
        assert t == [1 if (j == x + 1) else 0 for j in range(1 << n)]
    assert S.eq_evals_pointwise(r) == S.eq_evals(r)


def _columns(rng, T, trace_like):
    def col(kind):
        if not trace_like:
            return [rng.randrange(P) for _ in range(T)]
        if kind == "flag":
            return [rng.randrange(2) for _ in range(T)]
        if kind == "pc":
            return [0x80000000 + 4 * i for i in range(T)]
        return [rng.randrange(1 << 32) for _ in range(T)]
    kinds = {"unexpanded_pc": "pc", "pc": "pc", "is_virtual": "flag", "is_first_in_sequence": "flag", "is_noop": "flag",
             "left_is_rs1": "flag", "left_is_pc": "flag", "right_is_rs2": "flag", "right_is_imm": "flag"}
    names = set(S.ShiftPrefixSuffixProver.COLS) | set(S.InstructionInputProver.COLS) | set(S.RegistersPrefixSuffixProver.COLS)
    return {k: col(kinds.get(k, "value")) for k in sorted(names)}


@pytest.mark.parametrize("n,trace_like", [(2, False), (3, False), (4, True), (5, False), (6, True)])
def test_stage3_oracle_invariants(n, trace_like):
    rng = random.Random(100 + n)
    T = 1 << n
    cols = _columns(rng, T, trace_like)
    r_outer = [rng.randrange(P) for _ in range(n)]
    r_product = [rng.randrange(P) for _ in range(n)]
    g = rng.randrange(P)
    gp = [pow(g, i, P) for i in range(5)]
    ig, rg = rng.randrange(P), rng.randrange(P)
    shift = S.ShiftPrefixSuffixProver(cols, r_outer, r_product, gp)
    instr = S.InstructionInputProver(cols, r_outer, r_product, ig)
    reg = S.RegistersPrefixSuffixProver(cols, r_outer, rg)

    # SHIFT_INIT: grand_sum(P*Q) = direct_sum over eq+1 (:1141-1170)
    e1o, e1p = S.eq_plus_one_evals(r_outer), S.eq_plus_one_evals(r_product)
    direct = 0
    for j in range(T):
        v = (cols["unexpanded_pc"][j] + gp[1] * cols["pc"][j] + gp[2] * cols["is_virtual"][j] + gp[3] * cols["is_first_in_sequence"][j]) % P
        direct = (direct + e1o[j] * v + gp[4] * e1p[j] % P * ((1 - cols["is_noop"][j]) % P)) % P
    assert shift.input_claim() == direct
    eo = S.eq_evals(r_outer)
    assert reg.input_claim() == sum(eo[j] * (cols["rd_write_value"][j] + rg * cols["rs1_value"][j] + rg * rg % P * cols["rs2_value"][j])
                                    for j in range(T)) % P

    bc = [rng.randrange(P) for _ in range(3)]
    ch = [rng.randrange(P) for _ in range(n)]
    claims0 = [shift.input_claim(), instr.input_claim(), reg.input_claim()]
    rounds, final = S.run_stage3(shift, instr, reg, bc, ch)
    claims = claims0
    combined_claim = sum(c * b for c, b in zip(claims0, bc)) % P
    for e, r in zip(rounds, ch):
        assert (e["shift"][0] + e["shift"][1]) % P == claims[0]  # summed directly in phase 1, from the claim in phase 2
        assert (e["instr"][0] + e["instr"][1]) % P == claims[1]
        assert (e["reg"][0] + e["reg"][1]) % P == claims[2]
        assert (e["combined"][0] + e["combined"][1]) % P == combined_claim
        claims = [S._quad_at(e["shift"], r), evaluate_cubic_at_point(e["instr"], r), S._quad_at(e["reg"], r)]
        combined_claim = evaluate_cubic_at_point(e["combined"], r)
    assert claims == final
    # final-claim relations (:620-700); the challenges bind LSB first, the points are big-endian
    rev = ch[::-1]
    s = dict(zip(S.ShiftPrefixSuffixProver.COLS, shift.final_claims()))
    shift_val = (s["unexpanded_pc"] + gp[1] * s["pc"] + gp[2] * s["is_virtual"] + gp[3] * s["is_first_in_sequence"]) % P
    shift_expected = (S.eq_plus_one_mle(r_outer, rev) * shift_val + gp[4] * ((1 - s["is_noop"]) % P) % P * S.eq_plus_one_mle(r_product, rev)) % P
    assert shift_expected == final[0]
    i = dict(zip(S.InstructionInputProver.COLS, instr.final_claims()))
    left = (i["left_is_rs1"] * i["rs1_value"] + i["left_is_pc"] * i["unexpanded_pc"]) % P
    right = (i["right_is_rs2"] * i["rs2_value"] + i["right_is_imm"] * i["imm"]) % P
    assert (eq_mle(r_outer, rev) + ig * ig % P * eq_mle(r_product, rev)) % P * ((right + ig * left) % P) % P == final[1]
    rr = dict(zip(S.RegistersPrefixSuffixProver.COLS, reg.final_claims()))
    assert eq_mle(r_outer, rev) * ((rr["rd_write_value"] + rg * rr["rs1_value"] + rg * rg % P * rr["rs2_value"]) % P) % P == final[2]
    assert combined_claim == sum(c * b for c, b in zip(final, bc)) % P


def test_eq_plus_one_matches_the_reference_run():
    """Value-level vectors from the reference's captured run (logs/zolt.log -> tests/golden/zolt_log_kats.json): Stage 3's r_outer
    is the Stage 1 challenge vector, rounds 8..1 (SHIFT_INIT prints r_outer[0] and r_outer[last]), and SHIFT_INIT printed the
    first 8 bytes of eq+1(r_outer, k) for k = 0..4 and 255 and of eq(r_outer, 255) (stage3_prover.zig:1250-1290)."""
    import json
    from pathlib import Path

    d = json.loads((Path(__file__).resolve().parent / "golden" / "zolt_log_kats.json").read_text())
    ch = [int.from_bytes(bytes(c), "little") for c in d["stage1_challenges"]]
    assert len(ch) == 9
    r = [ch[i] for i in range(8, 0, -1)]
    assert list(r[0].to_bytes(32, "little")[:8]) == [70, 3, 78, 128, 115, 33, 214, 219]      # SHIFT_INIT: r_outer[0]
    assert list(r[-1].to_bytes(32, "little")[:8]) == [41, 25, 57, 228, 184, 23, 13, 80]     # SHIFT_INIT: r_outer[last]
    logged = {0: [0] * 8, 1: [166, 164, 37, 130, 206, 91, 156, 95], 2: [118, 194, 206, 228, 136, 232, 207, 27],
              3: [255, 73, 251, 187, 24, 103, 109, 214], 4: [9, 98, 82, 241, 126, 140, 187, 237], 255: [101, 82, 14, 229, 68, 52, 48, 82]}
    table = S.eq_plus_one_evals(r)
    for k, want in logged.items():
        assert list(S.eq_plus_one_mle(r, S._bits_be(k, 8)).to_bytes(32, "little")[:8]) == want
        assert list(table[k].to_bytes(32, "little")[:8]) == want
    assert list(eq_mle(r, S._bits_be(255, 8)).to_bytes(32, "little")[:8]) == [147, 167, 237, 58, 11, 177, 233, 250]
    # "eq(r, k-1) = eq+1(r, k)" as the run verified it for k = 1..4
    for k in range(1, 5):
        assert eq_mle(r, S._bits_be(k - 1, 8)) == table[k]
