"""Stage4GruenProver (src/zkvm/spartan/stage4_gruen_prover.zig) on the device against the loop-for-loop oracle
(oracle/stage4.py): every round polynomial, the bound tables' final values and the merged eq, over the phase configurations
the reference completes; sizes force each kernel family (streaming TMA rounds, small rounds, the resident tail) into use."""
import numpy as np
import pytest

import zolt_b200 as Z
from oracle import stage4 as S4
from oracle.bn254 import FR_MOD as P
from tests.util import fr_arr, fr_ints, rand_fr
from zolt_b200 import host as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built_lib):
    c = Z.Context(0)
    yield c
    c.close()


def _tables(seed, log_T, sparse):
    T = 1 << log_T
    n = S4.K * T
    if sparse:
        # trace-like: one-hot read/write addresses among 32 registers, small register values, increments of either sign
        rng = np.random.default_rng(seed)
        val = [int(v) for v in rng.integers(0, 1 << 32, n)]
        rs1, rs2, wa = [0] * n, [0] * n, [0] * n
        inc = [0] * T
        for j in range(T):
            rs1[int(rng.integers(0, 32)) * T + j] = 1
            if rng.integers(0, 2):
                rs2[int(rng.integers(0, 32)) * T + j] = 1
            if rng.integers(0, 4):
                wa[int(rng.integers(1, 32)) * T + j] = 1
                inc[j] = int(rng.integers(-(1 << 31), 1 << 31)) % P
        return val, wa, rs1, rs2, inc
    return (rand_fr(seed, n), rand_fr(seed + 1, n), rand_fr(seed + 2, n), rand_fr(seed + 3, n), rand_fr(seed + 4, T))


def _run(ctx, log_T, p1, seed, sparse=False, supply_claim=True):
    val, wa, rs1, rs2, inc = _tables(seed, log_T, sparse)
    gamma = rand_fr(seed + 5, 1)[0]
    r_cycle = rand_fr(seed + 6, log_T)
    challenges = rand_fr(seed + 7, S4.LOG_K + log_T)
    oracle = S4.Stage4Prover(val, wa, rs1, rs2, inc, gamma, r_cycle, p1, S4.LOG_K)
    claim0 = oracle.compute_input_claim()
    exp_evals, exp_last = oracle.run(challenges, claim0)
    bufs = [Z.DeviceBuffer.from_host(ctx, fr_arr(t)) for t in (val, wa, rs1, rs2, inc)]
    try:
        evals, _, finals, claims = H.stage4_run(ctx, log_T, *[b.ptr for b in bufs], fr_arr([gamma]), fr_arr(r_cycle), p1,
                                                input_claim=fr_arr([claim0]) if supply_claim else None, challenges=fr_arr(challenges))
    finally:
        for b in bufs:
            b.free()
    assert fr_ints(claims[0:1]) == [claim0]  # computeInputClaim (or the claim handed in)
    for rnd in range(len(challenges)):
        assert fr_ints(evals[rnd]) == exp_evals[rnd], f"log_T {log_T} phase1 {p1} round {rnd}"
    assert fr_ints(claims[1:2]) == [exp_last]
    fc = oracle.final_claims()
    assert fr_ints(finals) == [fc["val"], fc["rs1_ra"], fc["rs2_ra"], fc["rd_wa"], fc["inc"], oracle.merged_eq[0]]
    # the sumcheck's last claim is eq * body at the bound point (bindChallenge's closing check, :1192-1217)
    ra = (oracle.gamma * fc["rs1_ra"] + oracle.gamma_sq * fc["rs2_ra"]) % P
    assert oracle.merged_eq[0] * ((ra * fc["val"] + fc["rd_wa"] * (fc["val"] + fc["inc"])) % P) % P == exp_last


@pytest.mark.parametrize("log_T,p1", [(1, 1), (2, 1), (2, 2), (3, 1), (3, 2), (3, 3), (4, 2), (5, 2), (5, 5), (6, 3), (7, 1)])
def test_stage4_all_phases_small(ctx, log_T, p1):
    _run(ctx, log_T, p1, seed=4000 + 16 * log_T + p1)


def test_stage4_computes_its_input_claim(ctx):
    _run(ctx, 4, 2, seed=4500, supply_claim=False)


def test_stage4_trace_like_tables(ctx):
    _run(ctx, 6, 3, seed=4600, sparse=True, supply_claim=False)


def test_stage4_streaming_rounds(ctx):
    # K * T = 2^17 entries: phase-1 rounds 0-1 run on the streaming kernels (2^16 and 2^15 pairs), the rest on the small-round
    # kernel and the resident tail; phase 2 starts at 2^12 pairs
    _run(ctx, 10, 5, seed=4700)


def test_stage4_streaming_only(ctx):
    # no small-round kernel, no tail: every round of every phase on the streaming kernels (TMA boxes, ragged last chunk)
    ctx.set_tuning(Z.TUNE_SC_SMALL_PAIRS, 0)
    ctx.set_tuning(Z.TUNE_SC_TAIL_PAIRS, 0)
    try:
        _run(ctx, 6, 3, seed=4800)
    finally:
        ctx.set_tuning(Z.TUNE_SC_SMALL_PAIRS, 1 << 14)
        ctx.set_tuning(Z.TUNE_SC_TAIL_PAIRS, 1 << 12)


def test_stage4_prove_with_transcript(ctx):
    # Stage4GruenProver.prove: coefficients c0, c2, c3 (times the batching coefficient) into the transcript, challenge out;
    # replayed on the oracle with the squeezed challenges
    log_T, p1 = 5, 2
    val, wa, rs1, rs2, inc = _tables(4900, log_T, False)
    gamma, r_cycle = rand_fr(4905, 1)[0], rand_fr(4906, log_T)
    coeff = rand_fr(4907, 1)[0]
    bufs = [Z.DeviceBuffer.from_host(ctx, fr_arr(t)) for t in (val, wa, rs1, rs2, inc)]
    try:
        coeffs, chal, finals, claims = H.stage4_run(ctx, log_T, *[b.ptr for b in bufs], fr_arr([gamma]), fr_arr(r_cycle), p1,
                                                    batching_coeff=fr_arr([coeff]), domain=b"stage4-test")
    finally:
        for b in bufs:
            b.free()
    oracle = S4.Stage4Prover(val, wa, rs1, rs2, inc, gamma, r_cycle, p1, S4.LOG_K, batching_coeff=coeff)
    tr = H.Transcript(b"stage4-test")
    claim = oracle.compute_input_claim() * coeff % P
    assert fr_ints(claims[0:1]) == [oracle.compute_input_claim()]
    for rnd in range(oracle.num_rounds):
        c = oracle.round_coeffs(rnd, claim)
        assert fr_ints(coeffs[rnd]) == c, f"round {rnd}"
        for j in (0, 2, 3):
            tr.append_scalar(b"stage4_coeff", fr_arr([c[j] * coeff % P])[0])
        r = fr_ints(tr.challenge_scalar(b"stage4_r").reshape(1, 4))[0]
        assert fr_ints(chal[rnd:rnd + 1]) == [r]
        claim = S4.eval_coeffs(c, r) * coeff % P
        oracle.bind(rnd, r)
    fc = oracle.final_claims()
    assert fr_ints(finals[:5]) == [fc["val"], fc["rs1_ra"], fc["rs2_ra"], fc["rd_wa"], fc["inc"]]


def test_index_map_and_vec_ops(ctx):
    K, T = 8, 16
    src = rand_fr(5100, K * T)
    d_src = Z.DeviceBuffer.from_host(ctx, fr_arr(src))
    d_dst = Z.DeviceBuffer(ctx, 32 * K * T)
    lib = ctx.lib
    import ctypes as C
    Z._lib.check(lib.zolt_fr_index_map_dev(ctx.handle, C.c_void_p(d_dst.ptr), C.c_void_p(d_src.ptr), K * T, 1, K, T, K, T, 1), "index_map")
    assert fr_ints(d_dst.download((K * T, 4))) == [src[k * T + j] for j in range(T) for k in range(K)]  # transpose
    Z._lib.check(lib.zolt_fr_index_map_dev(ctx.handle, C.c_void_p(d_dst.ptr), C.c_void_p(d_src.ptr), K * T, 1, T, 1, 0, 0, 0), "index_map")
    assert fr_ints(d_dst.download((K * T, 4))) == [src[j] for _ in range(K) for j in range(T)]  # replicated rows
    Z._lib.check(lib.zolt_fr_index_map_dev(ctx.handle, C.c_void_p(d_dst.ptr), C.c_void_p(d_src.ptr), K * T, K, T, 1, 0, 0, 0), "index_map")
    assert fr_ints(d_dst.download((K * T, 4))) == [src[j] for j in range(T) for _ in range(K)]  # every element repeated
    a, b = rand_fr(5101, K * T), rand_fr(5102, K * T)
    s = rand_fr(5103, 2)
    d_a, d_b = Z.DeviceBuffer.from_host(ctx, fr_arr(a)), Z.DeviceBuffer.from_host(ctx, fr_arr(b))
    ctx.fr_vec_dev(Z.VEC_LIN2, d_dst.ptr, d_a.ptr, d_b.ptr, s=fr_arr(s), n=K * T)
    assert fr_ints(d_dst.download((K * T, 4))) == [(s[0] * x + s[1] * y) % P for x, y in zip(a, b)]
    for buf in (d_src, d_dst, d_a, d_b):
        buf.free()
