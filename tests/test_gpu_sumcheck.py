"""GPU parity: sumcheck round evaluations and bound tables, every round, vs the oracle.
Bit-exact (integer work).  Shapes follow the reference provers listed in oracle/sumcheck.py."""
import numpy as np
import pytest

import zolt_b200 as Z
from oracle import poly as OP
from oracle import sumcheck as OS
from tests.util import P, fr_arr, fr_ints, rand_fr

pytestmark = pytest.mark.gpu

COMBINERS = [
    ("sum", Z.SC_SUM, 1),
    ("prod2", Z.SC_PROD2, 2),
    ("prod3", Z.SC_PROD3, 3),
    ("eq_a_bmc", Z.SC_EQ_A_BMC, 4),
    ("eq_ab_minus_c", Z.SC_EQ_AB_MINUS_C, 4),
]


@pytest.fixture(params=["small_rounds", "streaming"])
def ctx(ctx, request):
    """Every sumcheck test runs twice: with the default dispatch (the latency-oriented small-round kernel below
    2^14 pairs) and with the streaming round kernels forced at every size."""
    ctx.set_tuning(Z.TUNE_SC_SMALL_PAIRS, (1 << 14) if request.param == "small_rounds" else 0)
    yield ctx
    ctx.set_tuning(Z.TUNE_SC_SMALL_PAIRS, 1 << 14)


def _run_and_compare(ctx, kind, comb, k, pairing, log_n, seed, n_extra=0):
    n = 1 << log_n
    tables = [rand_fr(seed * 100 + t, n) for t in range(k + n_extra)]
    challenges = rand_fr(seed * 100 + 99, log_n)
    exp_evals, exp_bound = OS.run_dense(kind, tables, pairing, challenges)
    inst = Z.SumcheckInstance(ctx, comb, pairing, [fr_arr(t) for t in tables[:k]], [fr_arr(t) for t in tables[k:]])
    try:
        for rnd, r in enumerate(challenges):
            got = fr_ints(inst.compute_round_poly())
            # oracle evaluates only the k combiner tables; extras are bound but not evaluated
            assert got == exp_evals[rnd], f"round {rnd} evals differ ({kind}, pairing {pairing}, n=2^{log_n})"
            inst.bind_challenge(fr_arr([r]))
            assert len(inst) == n >> (rnd + 1)
            for t in range(k + n_extra):
                assert fr_ints(inst.table(t)) == exp_bound[rnd][t], f"round {rnd} table {t} differs"
    finally:
        inst.free()


@pytest.mark.parametrize("kind,comb,k", COMBINERS)
@pytest.mark.parametrize("pairing", [Z.PAIR_ADJACENT, Z.PAIR_HALF])
@pytest.mark.parametrize("log_n", [1, 2, 3, 7, 11])
def test_dense_rounds_and_tables(ctx, kind, comb, k, pairing, log_n):
    _run_and_compare(ctx, kind, comb, k, pairing, log_n, seed=log_n + 17 * comb + pairing)


def test_fp64_pipe_bind_is_bit_identical(ctx):
    """ZOLT_TUNE_SC_BIND_F64: the fused passes multiply by the challenge on the FP64 pipe (16-bit x 32-bit limb products summed
    exactly in doubles) for the selected tables -- every round polynomial and every bound table must be unchanged."""
    try:
        for mask in (0xF, 0x5, 0x8):
            ctx.set_tuning(Z.TUNE_SC_BIND_F64, mask)
            _run_and_compare(ctx, "eq_ab_minus_c", Z.SC_EQ_AB_MINUS_C, 4, Z.PAIR_ADJACENT, 12, seed=41 + mask)
        ctx.set_tuning(Z.TUNE_SC_BIND_F64, 0x7)
        _run_and_compare(ctx, "prod3", Z.SC_PROD3, 3, Z.PAIR_ADJACENT, 16, seed=43)
        _run_and_compare(ctx, "eq_a_bmc", Z.SC_EQ_A_BMC, 4, Z.PAIR_ADJACENT, 9, seed=44, n_extra=1)
    finally:
        ctx.set_tuning(Z.TUNE_SC_BIND_F64, 0)


def test_shard_protocols_refuse_half_split_and_gruen(ctx):
    """Blocked hypercube shards are only meaningful for adjacent pairs and dense combiners: the C ABI must say
    so instead of returning plausible but wrong polynomials."""
    tabs = [fr_arr(rand_fr(70 + t, 16)) for t in range(2)]
    inst = Z.SumcheckInstance(ctx, Z.SC_PROD2, Z.PAIR_HALF, tabs)
    try:
        with pytest.raises(Exception):
            inst.set_sharded(True)
    finally:
        inst.free()
    inst = Z.SumcheckInstance(ctx, Z.SC_GRUEN_AB, Z.PAIR_ADJACENT, tabs, tau=fr_arr(rand_fr(72, 4)))
    try:
        with pytest.raises(Exception):
            inst.set_sharded(True)
    finally:
        inst.free()


def test_resident_tail_kernel(ctx):
    """The rounds of at most ZOLT_TUNE_SC_TAIL_PAIRS pairs run inside one resident cluster kernel that is fed challenges
    through mapped memory.  Same round polynomials and tables (a) when nothing interrupts it (whole proofs through the C++
    round loop), (b) when every round reads the tables back (each read joins the kernel, the next bind restarts it),
    (c) for two instances driven alternately on one context (a batched sumcheck), (d) when the host pauses longer than
    the kernel's patience (the command is then served by a regular launch), and (e) with the tail switched off."""
    import time

    from oracle import cpu_ref as R
    from zolt_b200 import host as H

    def prove(kind, comb, k, log_n, pairing=Z.PAIR_ADJACENT, seed=0):
        tabs = [fr_arr(rand_fr(9000 + seed + t, 1 << log_n)) for t in range(k)]
        bufs = [Z.DeviceBuffer.from_host(ctx, t) for t in tabs]
        try:
            got = H.sumcheck_prove_dev(ctx, comb, pairing, [b.ptr for b in bufs], log_n)
        finally:
            for b in bufs:
                b.free()
        want = R.sc_prove(kind, pairing, tabs)
        ne = {Z.SC_SUM: 2, Z.SC_PROD2: 3}.get(comb, 4)
        assert (got[1] == want[1]).all() and (got[0][:, :ne] == want[0][:, :ne]).all(), (kind, log_n)
        assert all((got[2][t] == want[2][t][0]).all() for t in range(k)), (kind, log_n)

    try:
        for pairs in (1 << 12, 1 << 14, 64, 0):
            ctx.set_tuning(Z.TUNE_SC_TAIL_PAIRS, pairs)
            prove("eq_ab_minus_c", Z.SC_EQ_AB_MINUS_C, 4, 15, seed=pairs)
            prove("prod2", Z.SC_PROD2, 2, 13, pairing=Z.PAIR_HALF, seed=pairs + 1)
            prove("sum", Z.SC_SUM, 1, 9, seed=pairs + 2)
            prove("prod3", Z.SC_PROD3, 3, 3, seed=pairs + 3)
        ctx.set_tuning(Z.TUNE_SC_TAIL_PAIRS, 1 << 12)
        # (b) tables read back every round, (with an extra bound-only table)
        _run_and_compare(ctx, "eq_a_bmc", Z.SC_EQ_A_BMC, 4, Z.PAIR_ADJACENT, 10, seed=61, n_extra=1)
        # (c) two instances on one context, rounds interleaved
        log_n = 11
        ta = [rand_fr(9100 + t, 1 << log_n) for t in range(3)]
        tb = [rand_fr(9200 + t, 1 << log_n) for t in range(2)]
        ch = rand_fr(9300, log_n)
        ea, _ = OS.run_dense("prod3", ta, Z.PAIR_ADJACENT, ch)
        eb, _ = OS.run_dense("prod2", tb, Z.PAIR_ADJACENT, ch)
        ia = Z.SumcheckInstance(ctx, Z.SC_PROD3, Z.PAIR_ADJACENT, [fr_arr(t) for t in ta])
        ib = Z.SumcheckInstance(ctx, Z.SC_PROD2, Z.PAIR_ADJACENT, [fr_arr(t) for t in tb])
        try:
            for rnd, r in enumerate(ch):
                assert fr_ints(ia.compute_round_poly()) == ea[rnd] and fr_ints(ib.compute_round_poly()) == eb[rnd], rnd
                ia.bind_challenge(fr_arr([r]))
                ib.bind_challenge(fr_arr([r]))
        finally:
            ia.free()
            ib.free()
        # (d) a host slower than the kernel's patience
        ctx.set_tuning(Z.TUNE_SC_TAIL_TIMEOUT_MS, 50)
        tabs = [rand_fr(9400 + t, 1 << 8) for t in range(4)]
        ch = rand_fr(9500, 8)
        exp, bound = OS.run_dense("eq_ab_minus_c", tabs, Z.PAIR_ADJACENT, ch)
        inst = Z.SumcheckInstance(ctx, Z.SC_EQ_AB_MINUS_C, Z.PAIR_ADJACENT, [fr_arr(t) for t in tabs])
        try:
            for rnd, r in enumerate(ch):
                assert fr_ints(inst.compute_round_poly()) == exp[rnd], rnd
                if rnd in (2, 5, len(ch) - 1):  # ... including before the LAST bind, which nobody waits for
                    time.sleep(0.25)
                inst.bind_challenge(fr_arr([r]))
            for t in range(4):
                assert fr_ints(inst.table(t)) == bound[-1][t]
        finally:
            inst.free()
    finally:
        ctx.set_tuning(Z.TUNE_SC_TAIL_TIMEOUT_MS, 0)
        ctx.set_tuning(Z.TUNE_SC_TAIL_PAIRS, 1 << 12)


def test_dispatch_crossover(ctx):
    # 2^16 entries: rounds 0-1 stream, the rest run on the small-round kernel (several chunks per block)
    _run_and_compare(ctx, "eq_ab_minus_c", Z.SC_EQ_AB_MINUS_C, 4, Z.PAIR_ADJACENT, 16, seed=31)
    _run_and_compare(ctx, "prod2", Z.SC_PROD2, 2, Z.PAIR_HALF, 16, seed=32)


def test_extra_tables_are_bound_in_lockstep(ctx):
    # OutputSumcheckProver binds val_init too (output_check.zig:454-476)
    _run_and_compare(ctx, "eq_a_bmc", Z.SC_EQ_A_BMC, 4, Z.PAIR_ADJACENT, 9, seed=5, n_extra=1)


@pytest.mark.parametrize("log_n", [1, 2, 3, 4, 8, 11])
@pytest.mark.parametrize("with_c", [False, True])
def test_gruen_rounds(ctx, log_n, with_c):
    n = 1 << log_n
    k = 3 if with_c else 2
    tables = [rand_fr(900 + log_n * 10 + t, n) for t in range(k)]
    tau = rand_fr(777 + log_n, log_n)
    challenges = rand_fr(778 + log_n, log_n)
    exp_evals, exp_bound = OS.run_gruen(tables, tau, challenges, with_c)
    comb = Z.SC_GRUEN_AB_MINUS_C if with_c else Z.SC_GRUEN_AB
    inst = Z.SumcheckInstance(ctx, comb, Z.PAIR_ADJACENT, [fr_arr(t) for t in tables], tau=fr_arr(tau))
    try:
        for rnd, r in enumerate(challenges):
            assert fr_ints(inst.compute_round_poly()) == exp_evals[rnd], f"round {rnd}"
            inst.bind_challenge(fr_arr([r]))
            for t in range(k):
                assert fr_ints(inst.table(t)) == exp_bound[rnd][t]
    finally:
        inst.free()


def test_reference_unit_vectors(ctx):
    # bindLow numeric examples (src/poly/mod.zig:816-888)
    p = Z.DensePolynomial(ctx, fr_arr([1, 2, 3, 4]))
    p.bind_low(fr_arr([3]))
    assert fr_ints(p.evaluations()) == [4, 6]
    p.free()
    p = Z.DensePolynomial(ctx, fr_arr([10, 20, 30, 40, 50, 60, 70, 80]))
    p.bind_low(fr_arr([5]))
    assert fr_ints(p.evaluations()) == [60, 80, 100, 120]
    p.free()
    # half-split sumcheck rounds on [1,2,3,4] (src/subprotocols/mod.zig:379-436): g(0)=3, g(1)=7, r=2 -> [5,6]
    inst = Z.SumcheckInstance(ctx, Z.SC_SUM, Z.PAIR_HALF, [fr_arr([1, 2, 3, 4])])
    assert fr_ints(inst.compute_round_poly()) == [3, 7]
    inst.bind_challenge(fr_arr([2]))
    assert fr_ints(inst.table(0)) == [5, 6]
    inst.free()
    # total 36 over 8 values (src/subprotocols/mod.zig:441-461)
    p = Z.DensePolynomial(ctx, fr_arr([1, 2, 3, 4, 5, 6, 7, 8]))
    assert fr_ints(p.sum()) == [36]
    p.bind_first(fr_arr([2]))  # (1-2)*lo + 2*hi
    assert fr_ints(p.evaluations()) == [(2 * (5 + i) - (1 + i)) % P for i in range(4)]
    p.free()


@pytest.mark.parametrize("n_vars", [0, 1, 2, 5, 12])
def test_eq_evals(ctx, n_vars):
    r = rand_fr(4242 + n_vars, n_vars)
    got = fr_ints(Z.EqPolynomial.evals(ctx, fr_arr(r) if n_vars else np.empty((0, 4), dtype=np.uint64)))
    assert got == OP.eq_evals(r)
    assert sum(got) % P == 1  # partition of unity (src/poly/mod.zig:689-753)
    scale = rand_fr(1, 1)
    got = fr_ints(Z.EqPolynomial.evals(ctx, fr_arr(r) if n_vars else np.empty((0, 4), dtype=np.uint64), fr_arr(scale)))
    assert got == OP.eq_evals(r, scale[0])


def test_call_sequence_errors(ctx):
    inst = Z.SumcheckInstance(ctx, Z.SC_SUM, Z.PAIR_ADJACENT, [fr_arr([1, 2])])
    inst.compute_round_poly()
    inst.bind_challenge(fr_arr([9]))
    with pytest.raises(Z.ZoltGpuError):
        inst.bind_challenge(fr_arr([9]))  # all variables bound (ZOLT_ERR_STATE)
    with pytest.raises(Z.ZoltGpuError):
        inst.compute_round_poly()
    inst.free()


def test_device_table_construction_and_spartan_stage1(ctx):
    """JoltSpartanInterface builds combined = eq * (Az * Bz - Cz) (jolt_r1cs.zig:352-360) and runs
    a linear sumcheck over it with adjacent pairs (:436-477).  Here the eq table, the combined table and the rounds are all
    produced on the device; nothing but tau, the challenges and the round sums crosses the bus."""
    log_n = 10
    n = 1 << log_n
    tau = rand_fr(61, log_n)
    az, bz, cz = rand_fr(62, n), rand_fr(63, n), rand_fr(64, n)
    eq = OP.eq_evals(tau)
    want = [e * (a * b - c) % P for e, a, b, c in zip(eq, az, bz, cz)]
    d_eq = Z.DeviceBuffer(ctx, n * 32)
    ctx_lib = ctx.lib
    import ctypes as C

    from zolt_b200._lib import check
    check(ctx_lib.zolt_eq_evals(ctx.handle, fr_arr(tau).ctypes.data_as(C.c_void_p), log_n, None, C.c_void_p(d_eq.ptr)), "zolt_eq_evals")
    d_a, d_b, d_c = (Z.DeviceBuffer.from_host(ctx, fr_arr(t)) for t in (az, bz, cz))
    d_out = Z.DeviceBuffer(ctx, n * 32)
    ctx.fr_vec_dev(Z.VEC_EQ_AB_MINUS_C, d_out.ptr, d_a.ptr, d_b.ptr, d_c.ptr, d_eq.ptr, n=n)
    assert fr_ints(d_out.download((n, 4))) == want
    # out = a o b, then out += s * c
    ctx.fr_vec_dev(Z.VEC_MUL, d_out.ptr, d_a.ptr, d_b.ptr, n=n)
    s = rand_fr(65, 1)[0]
    ctx.fr_vec_dev(Z.VEC_AXPY, d_out.ptr, d_c.ptr, s=fr_arr([s]), n=n)
    assert fr_ints(d_out.download((n, 4))) == [(a * b + s * c) % P for a, b, c in zip(az, bz, cz)]
    # stage 1 rounds over the combined table straight from device memory
    ctx.fr_vec_dev(Z.VEC_EQ_AB_MINUS_C, d_out.ptr, d_a.ptr, d_b.ptr, d_c.ptr, d_eq.ptr, n=n)
    challenges = rand_fr(66, log_n)
    exp_evals, exp_bound = OS.run_dense("sum", [want], Z.PAIR_ADJACENT, challenges)
    inst = Z.SumcheckInstance.from_device(ctx, Z.SC_SUM, Z.PAIR_ADJACENT, [d_out.ptr], log_n)
    for rnd, r in enumerate(challenges):
        assert fr_ints(inst.compute_round_poly()) == exp_evals[rnd]
        inst.bind_challenge(fr_arr([r]))
        assert fr_ints(inst.table(0)) == exp_bound[rnd][0]
    inst.free()
    for b in (d_eq, d_a, d_b, d_c, d_out):
        b.free()


@pytest.mark.parametrize("log_k", [1, 4, 10])
def test_raf_evaluation_prover_shape(ctx, log_k):
    """S3: RafEvaluationProver (src/zkvm/ram/raf_checking.zig:335-409) keeps `unmap` implicit; on the device it is the
    PROD2 combiner over (ra, unmap table) with the prover's running claim supplying s(1).  The unmap table
    start_address + 8 k is generated on the device (ZOLT_VEC_AFFINE_INDEX)."""
    k = 1 << log_k
    ra = rand_fr(1200 + log_k, k)
    start = 0x8000_0000
    challenges = rand_fr(1300 + log_k, log_k)
    claim = sum(a * (start + 8 * i) for i, a in enumerate(ra)) % P  # computeInitialClaim (:316-324)
    exp_evals, exp_tables, exp_final = OS.run_raf(ra, start, challenges, claim)
    d_unmap = Z.DeviceBuffer(ctx, k * 32)
    ctx.fr_vec_dev(Z.VEC_AFFINE_INDEX, d_unmap.ptr, s=fr_arr([start, 8]), n=k)
    assert fr_ints(d_unmap.download((k, 4))) == [start + 8 * i for i in range(k)]
    d_ra = Z.DeviceBuffer.from_host(ctx, fr_arr(ra))
    inst = Z.SumcheckInstance.from_device(ctx, Z.SC_PROD2, Z.PAIR_ADJACENT, [d_ra.ptr, d_unmap.ptr], log_k)
    inst.set_claim(fr_arr([claim]))
    for rnd, r in enumerate(challenges):
        assert fr_ints(inst.compute_round_poly()) == exp_evals[rnd], f"round {rnd}"
        inst.bind_challenge(fr_arr([r]))
        assert fr_ints(inst.table(0)) == exp_tables[rnd]
    inst.free()
    d_ra.free()
    d_unmap.free()


def test_serialised_launches_disable_the_resident_tail(built_lib, tmp_path):
    """Under a profiler (ncu) or CUDA_LAUNCH_BLOCKING=1 a launch returns only when the kernel has retired, so a kernel that
    waits for the host can only time out: the library must notice (first launch), fall back to one launch per round for the
    context, and still produce the same proof."""
    import os
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    script = tmp_path / "blocking.py"
    script.write_text(f"""
import sys, time
sys.path.insert(0, {str(root)!r})
import numpy as np
import zolt_b200 as Z
from zolt_b200 import host as H
from oracle import cpu_ref as R
from tests.util import fr_arr, rand_fr
ctx = Z.Context(0)
tabs = [fr_arr(rand_fr(9600 + t, 1 << 13)) for t in range(4)]
bufs = [Z.DeviceBuffer.from_host(ctx, t) for t in tabs]
t0 = time.time()
for _ in range(2):
    got = H.sumcheck_prove_dev(ctx, Z.SC_EQ_AB_MINUS_C, Z.PAIR_ADJACENT, [b.ptr for b in bufs], 13)
want = R.sc_prove("eq_ab_minus_c", 0, tabs)
assert (got[0] == want[0]).all() and (got[1] == want[1]).all()
assert all((got[2][t] == want[2][t][0]).all() for t in range(4))
assert time.time() - t0 < 5.0, "every round paid a time-out"
print("BLOCKING_OK")
""")
    res = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=300, env=dict(os.environ, CUDA_LAUNCH_BLOCKING="1"))
    assert res.returncode == 0 and "BLOCKING_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-3000:]
    assert "resident sumcheck tail disabled" in res.stderr
