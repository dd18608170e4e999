"""Stage 3 (src/zkvm/spartan/stage3_prover.zig) on the device against the loop-for-loop oracle (oracle/stage3.py): the eq+1
table, the Q-buffer fold, the w * (a b + c d) combiner, and the three provers (Shift / InstructionInput / Registers) run
together like generateStage3Proof's loop -- every prover's evaluations in every round, the batched cubic, the running claims
and the 16 final claims; sizes force the streaming rounds, the small-round kernel and the resident tail into use."""
import random

import pytest

import zolt_b200 as Z
from oracle import stage3 as S3
from oracle.bn254 import FR_MOD as P
from tests.test_oracle_stage3 import _columns
from tests.util import fr_arr, fr_ints, rand_fr
from zolt_b200 import host as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built_lib):
    c = Z.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("n", [1, 2, 3, 5, 9, 13])
def test_eq_plus_one_evals(ctx, n):
    r = rand_fr(300 + n, n)
    out = Z.DeviceBuffer(ctx, 32 << n)
    try:
        ctx.eq_plus_one_evals_dev(fr_arr(r), out.ptr)
        got = fr_ints(out.download((1 << n, 4)))
    finally:
        out.free()
    if n <= 9:
        assert got == S3.eq_plus_one_evals(r)
    else:  # the literal O(n^2)-per-entry table takes too long: spot entries + the successor identity on a boolean point
        for j in (0, 1, 2, 3, (1 << n) - 1, 1 << (n - 1), 0x1234 & ((1 << n) - 1)):
            assert got[j] == S3.eq_plus_one_mle(r, S3._bits_be(j, n))
        assert sum(got) % P == (1 - _prod(r)) % P  # sum_j eq+1(r, j) = 1 - prod r_i
    # boolean point x: indicator of x + 1
    x = (0b1011011 % ((1 << n) - 1)) if n > 1 else 0
    xb = fr_arr(S3._bits_be(x, n))
    out = Z.DeviceBuffer(ctx, 32 << n)
    try:
        ctx.eq_plus_one_evals_dev(xb, out.ptr)
        got = fr_ints(out.download((1 << n, 4)))
    finally:
        out.free()
    assert got == [1 if j == x + 1 else 0 for j in range(1 << n)]


def _prod(r):
    m = 1
    for x in r:
        m = m * x % P
    return m


@pytest.mark.parametrize("n_lo,n_hi,scaled", [(1, 1, False), (2, 8, True), (16, 16, False), (32, 5, True), (48, 64, False), (256, 512, True),
                                               (4096, 64, False)])
def test_fold_high(ctx, n_lo, n_hi, scaled):
    rng = random.Random(n_lo * 131 + n_hi)
    small = n_lo * n_hi > 1 << 14  # keep the Python side cheap on the big shape: 32-bit entries, full-width weights
    tab = [rng.randrange(1 << 32) if small else rng.randrange(P) for _ in range(n_lo * n_hi)]
    w = [rng.randrange(P) for _ in range(n_hi)]
    scale = rng.randrange(P) if scaled else None
    bufs = [Z.DeviceBuffer.from_host(ctx, fr_arr(t)) for t in (tab, w)]
    out = Z.DeviceBuffer(ctx, 32 * n_lo)
    try:
        ctx.fr_fold_high_dev(out.ptr, bufs[0].ptr, bufs[1].ptr, n_lo, n_hi, fr_arr([scale]) if scaled else None)
        got = fr_ints(out.download((n_lo, 4)))
    finally:
        for b in bufs + [out]:
            b.free()
    exp = [sum(tab[x + h * n_lo] * w[h] for h in range(n_hi)) % P * (scale if scaled else 1) % P for x in range(n_lo)]
    assert got == exp


def _dot2_oracle(tables, challenges):
    """w (a b + c d) sampled at X = 0, 2, 3 per round, tables bound low-to-high."""
    t = [list(x) for x in tables]
    evals, bound = [], []
    for r in challenges:
        e = [0, 0, 0]
        for j in range(len(t[0]) // 2):
            lo = [x[2 * j] for x in t]
            d = [(x[2 * j + 1] - x[2 * j]) % P for x in t]
            for slot, pt in enumerate((0, 2, 3)):
                v = [(a + pt * b) % P for a, b in zip(lo, d)]
                e[slot] = (e[slot] + v[0] * (v[1] * v[2] + v[3] * v[4])) % P
        evals.append(e)
        t = [S3._bind(x, r) for x in t]
        bound.append([list(x) for x in t])
    return evals, bound


@pytest.mark.parametrize("log_n", [1, 2, 3, 6, 10, 13, 16])
def test_eq_dot2_rounds(ctx, log_n):
    n = 1 << log_n
    tables = [rand_fr(4100 + 7 * log_n + k, n) for k in range(5)]
    challenges = rand_fr(4200 + log_n, log_n)
    exp_evals, exp_bound = _dot2_oracle(tables, challenges)
    inst = Z.SumcheckInstance(ctx, Z.SC_EQ_DOT2, Z.PAIR_ADJACENT, [fr_arr(t) for t in tables])
    try:
        for rnd, r in enumerate(challenges):
            assert fr_ints(inst.compute_round_poly()) == exp_evals[rnd], f"round {rnd}"
            inst.bind_challenge(fr_arr([r]))
            if rnd in (0, 1, log_n - 1):
                for k in range(5):
                    assert fr_ints(inst.table(k)) == exp_bound[rnd][k]
    finally:
        inst.free()


def test_eq_dot2_streaming_kernels_only(ctx):
    """the same rounds with the small-round kernel and the resident tail switched off: every round on sc_round_kernel"""
    ctx.set_tuning(Z.TUNE_SC_SMALL_PAIRS, 0)
    ctx.set_tuning(Z.TUNE_SC_TAIL_PAIRS, 0)
    try:
        test_eq_dot2_rounds(ctx, 10)
    finally:
        ctx.set_tuning(Z.TUNE_SC_SMALL_PAIRS, 1 << 14)
        ctx.set_tuning(Z.TUNE_SC_TAIL_PAIRS, 1 << 12)


def test_eq_dot2_refuses_half_pairing(ctx):
    with pytest.raises(Exception):
        Z.SumcheckInstance(ctx, Z.SC_EQ_DOT2, Z.PAIR_HALF, [fr_arr(rand_fr(k, 8)) for k in range(5)])


def _run_stage3(ctx, n, trace_like, seed):
    rng = random.Random(seed)
    T = 1 << n
    cols = _columns(rng, T, trace_like)
    r_outer = [rng.randrange(P) for _ in range(n)]
    r_product = [rng.randrange(P) for _ in range(n)]
    g = rng.randrange(P)
    gp = [pow(g, i, P) for i in range(5)]
    ig, rg = rng.randrange(P), rng.randrange(P)
    bc = [rng.randrange(P) for _ in range(3)]
    ch = [rng.randrange(P) for _ in range(n)]
    shift = S3.ShiftPrefixSuffixProver(cols, r_outer, r_product, gp)
    instr = S3.InstructionInputProver(cols, r_outer, r_product, ig)
    reg = S3.RegistersPrefixSuffixProver(cols, r_outer, rg)
    claims0 = [shift.input_claim(), instr.input_claim(), reg.input_claim()]
    exp_rounds, exp_final = S3.run_stage3(shift, instr, reg, bc, ch)
    exp_finals = shift.final_claims() + instr.final_claims() + reg.final_claims()

    bufs = {}
    d_cols = []
    for name in H.STAGE3_COLUMNS:
        col = name.split(".")[1]
        if col not in bufs:
            bufs[col] = Z.DeviceBuffer.from_host(ctx, fr_arr(cols[col]))
        d_cols.append(bufs[col].ptr)
    try:
        evals, claims, finals = H.stage3_run(ctx, n, d_cols, fr_arr(r_outer), fr_arr(r_product), fr_arr(gp + [ig, rg]), fr_arr(bc), fr_arr(ch))
    finally:
        for b in bufs.values():
            b.free()
    assert fr_ints(claims[:3]) == claims0
    for rnd in range(n):
        e = fr_ints(evals[rnd])
        assert e[0:3] == exp_rounds[rnd]["shift"], f"shift round {rnd}"
        assert e[3:7] == exp_rounds[rnd]["instr"], f"instr round {rnd}"
        assert e[7:10] == exp_rounds[rnd]["reg"], f"reg round {rnd}"
        assert e[10:14] == exp_rounds[rnd]["combined"], f"combined round {rnd}"
    assert fr_ints(claims[3:]) == exp_final
    assert fr_ints(finals) == exp_finals


@pytest.mark.parametrize("n,trace_like", [(2, False), (3, False), (4, True), (5, False), (8, True), (9, False), (12, True)])
def test_stage3_provers_match_oracle(ctx, n, trace_like):
    _run_stage3(ctx, n, trace_like, 500 + n)


def test_stage3_streaming_kernels_only(ctx):
    ctx.set_tuning(Z.TUNE_SC_SMALL_PAIRS, 0)
    ctx.set_tuning(Z.TUNE_SC_TAIL_PAIRS, 0)
    try:
        _run_stage3(ctx, 8, False, 77)
    finally:
        ctx.set_tuning(Z.TUNE_SC_SMALL_PAIRS, 1 << 14)
        ctx.set_tuning(Z.TUNE_SC_TAIL_PAIRS, 1 << 12)


def test_device_eq_plus_one_table_matches_the_reference_run(ctx):
    """The device table for r_outer of the reference's captured run (its Stage 1 challenges, rounds 8..1) against the values
    that run printed: eq+1(r_outer, k), first 8 bytes, k = 0..4 and 255 (logs/zolt.log SHIFT_INIT -> tests/golden)."""
    import json
    from pathlib import Path

    d = json.loads((Path(__file__).resolve().parent / "golden" / "zolt_log_kats.json").read_text())
    ch = [int.from_bytes(bytes(c), "little") for c in d["stage1_challenges"]]
    r = [ch[i] for i in range(8, 0, -1)]
    out = Z.DeviceBuffer(ctx, 32 << 8)
    try:
        ctx.eq_plus_one_evals_dev(fr_arr(r), out.ptr)
        got = fr_ints(out.download((256, 4)))
    finally:
        out.free()
    logged = {0: [0] * 8, 1: [166, 164, 37, 130, 206, 91, 156, 95], 2: [118, 194, 206, 228, 136, 232, 207, 27],
              3: [255, 73, 251, 187, 24, 103, 109, 214], 4: [9, 98, 82, 241, 126, 140, 187, 237], 255: [101, 82, 14, 229, 68, 52, 48, 82]}
    for k, want in logged.items():
        assert list(got[k].to_bytes(32, "little")[:8]) == want
