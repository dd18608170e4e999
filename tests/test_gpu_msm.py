"""GPU parity: BN254 G1 MSM through the C ABI vs the oracle (literal Pippenger restatement, closed
form (sum s_i k_i) G, and the mock-SRS commit KAT).  Bit-exact affine Montgomery limbs + infinity flag.
Reference: src/msm/mod.zig:355-438, src/poly/commitment/mod.zig:174-255."""
import numpy as np
import pytest

import zolt_b200 as Z
from oracle import bn254 as B
from tests.util import P, fr_arr, point_from_out, rand_fr

pytestmark = pytest.mark.gpu


def _bases_from_logs(ctx, logs):
    """bases[i] = logs[i] * G on the GPU (fixed-base kernel), checked elsewhere against the oracle."""
    return Z.Bases.fixed_base_mul(ctx, fr_arr(logs))


def test_fixed_base_mul_matches_oracle(ctx):
    logs = [0, 1, 2, 3, P - 1, P - 2, 255, 256, 0x12345678, (1 << 253) + 12345] + rand_fr(21, 54)
    bases = _bases_from_logs(ctx, logs)
    xy, inf = bases.download()
    got = B.arrays_to_points(xy, inf)
    exp = [B.g1_mul(B.G1_GEN, k) for k in logs]
    assert got == exp
    assert all(B.g1_is_on_curve(p) for p in got)
    bases.free()


def test_hyperkzg_setup_is_powers_of_tau(ctx):
    # 'hyperkzg srs has correct tau relationship' (commitment/mod.zig:1260-1299): powers[1] = tau * G
    params = Z.HyperKZG.setup(ctx, 40)
    xy, inf = params.download()
    assert B.arrays_to_points(xy, inf) == B.hyperkzg_setup(40)
    params.free()


def test_commit_kat_1234(ctx):
    # SURVEY 8c.3 derived KAT: commit([1,2,3,4]) with the mock SRS
    params = Z.HyperKZG.setup(ctx, 4)
    xy, inf = Z.HyperKZG.commit(params, fr_arr([1, 2, 3, 4]))
    assert inf == 0
    assert point_from_out(xy, inf) == (
        0x07AEB450B7B66FDF44A37824224849B29C1ED87D705324299A17C8FB790F75CE,
        0x166808E4F5E0101BE322704903A3156ABBE3AD2C7A8992563F704CE020B05D14,
    )
    assert [int(v) for v in xy[:4]] == [0x34686E9C2DC1A390, 0x6E093777DB13CF10, 0xAE87848108CE376F, 0x23A755E5AA7A0FD3]
    assert B.hyperkzg_commit_closed_form([1, 2, 3, 4]) == point_from_out(xy, inf)
    params.free()


@pytest.mark.parametrize("n", [0, 1, 2, 7, 8, 9, 31, 100, 513, 3000])
def test_msm_random_vs_oracle(ctx, n):
    logs = [k or 1 for k in rand_fr(1000 + n, n)]
    scalars = rand_fr(2000 + n, n)
    bases = _bases_from_logs(ctx, logs)
    xy, inf = Z.MSM.compute(bases, fr_arr(scalars) if n else np.empty((0, 4), dtype=np.uint64))
    got = point_from_out(xy, inf)
    assert got == B.msm_closed_form(logs, scalars)
    if n <= 513:
        pts = B.arrays_to_points(*bases.download()) if n else []
        assert got == B.msm(pts, scalars)  # literal reference algorithm (naive < 8, Pippenger otherwise)
    if got is None:
        assert inf == 1 and not xy.any()  # identity is (0, 0, infinity) (src/msm/mod.zig:24-30)
    bases.free()


def test_msm_edge_cases(ctx):
    n = 64
    logs = rand_fr(5, n)
    bases = _bases_from_logs(ctx, logs)
    # zero scalars -> identity
    xy, inf = Z.MSM.compute(bases, fr_arr([0] * n))
    assert inf == 1 and not xy.any()
    # 1 * P = P (src/msm/mod.zig:853-873 style)
    s = [0] * n
    s[3] = 1
    xy, inf = Z.MSM.compute(bases, fr_arr(s))
    assert point_from_out(xy, inf) == B.g1_mul(B.G1_GEN, logs[3])
    # scalar r-1 == -1
    s[3] = P - 1
    xy, inf = Z.MSM.compute(bases, fr_arr(s))
    assert point_from_out(xy, inf) == B.g1_neg(B.g1_mul(B.G1_GEN, logs[3]))
    bases.free()


def test_msm_equal_and_opposite_points(ctx):
    # every base equal: bucket accumulation must take the doubling path (P + P)
    n = 300
    bases = _bases_from_logs(ctx, [7] * n)
    scalars = [5] * n
    xy, inf = Z.MSM.compute(bases, fr_arr(scalars))
    assert point_from_out(xy, inf) == B.g1_mul(B.G1_GEN, 7 * 5 * n)
    # pairs (P, -P) with equal scalars cancel exactly: P + (-P) = identity
    logs = []
    for k in rand_fr(8, n // 2):
        logs += [k, P - k]
    bases2 = _bases_from_logs(ctx, logs)
    sc = []
    for s in rand_fr(9, n // 2):
        sc += [s, s]
    xy, inf = Z.MSM.compute(bases2, fr_arr(sc))
    assert inf == 1 and not xy.any()
    # all scalars equal (one bucket per window holds every point)
    xy, inf = Z.MSM.compute(bases2, fr_arr([12345] * n))
    assert inf == 1
    bases.free()
    bases2.free()


def test_msm_infinity_bases_are_skipped(ctx):
    n = 40
    logs = rand_fr(31, n)
    pts = [B.g1_mul(B.G1_GEN, k) for k in logs]
    pts[0] = None
    pts[17] = None
    xy, inf = B.points_to_arrays(pts)
    bases = Z.Bases.upload(ctx, xy, inf)
    scalars = rand_fr(32, n)
    out, oinf = Z.MSM.compute(bases, fr_arr(scalars))
    exp_logs = [0 if p is None else k for p, k in zip(pts, logs)]
    assert point_from_out(out, oinf) == B.msm_closed_form(exp_logs, scalars)
    bases.free()


def test_msm_bench_zig_distribution(ctx):
    # src/bench.zig:243-287: scalars 7i+13, bases (i+1) G -- skewed digits (only low windows used)
    for n in (16, 64, 256, 4096):
        logs = [i + 1 for i in range(n)]
        scalars = [7 * i + 13 for i in range(n)]
        bases = _bases_from_logs(ctx, logs)
        xy, inf = Z.MSM.compute(bases, fr_arr(scalars))
        assert point_from_out(xy, inf) == B.msm_closed_form(logs, scalars)
        bases.free()


def test_msm_offset_and_batch(ctx):
    n, k = 200, 5
    logs = rand_fr(77, n + 10)
    bases = _bases_from_logs(ctx, logs)
    scalars = rand_fr(78, n)
    xy, inf = Z.MSM.compute(bases, fr_arr(scalars), offset=10)
    assert point_from_out(xy, inf) == B.msm_closed_form(logs[10:], scalars)
    batch = [rand_fr(100 + j, n) for j in range(k)]
    flat = np.concatenate([fr_arr(b) for b in batch])
    oxy, oinf = Z.MSM.compute_batch(bases, flat, k, n)
    for j in range(k):
        assert point_from_out(oxy[j], oinf[j]) == B.msm_closed_form(logs[:n], batch[j])
    bases.free()


@pytest.mark.parametrize("log_n", [14, 16])
def test_msm_large_closed_form(ctx, log_n):
    # BASELINE config 1 (2^16): GPU == closed form (sum s_i k_i) G
    n = 1 << log_n
    logs = rand_fr(0x5A170002, n)
    scalars = rand_fr(0x5A170001, n)
    bases = _bases_from_logs(ctx, logs)
    xy, inf = Z.MSM.compute(bases, fr_arr(scalars))
    assert point_from_out(xy, inf) == B.msm_closed_form(logs, scalars)
    bases.free()


@pytest.mark.parametrize("n,c", [(64, 8), (300, 9), (1000, 10), (4096, 13), (1 << 14, 18), (1 << 16, 22)])
def test_msm_precomputed_levels(ctx, n, c):
    """zolt_msm_bases_precompute: the window-level table changes the algorithm (one bucket set, W*n table
    points), never the result.  Small windows force the table path at small n (threshold 2^(c-6))."""
    logs = [k or 1 for k in rand_fr(4000 + n, n)]
    scalars = rand_fr(5000 + n, n)
    scalars[0], scalars[1] = P - 1, 0  # top-window carry, skipped digit
    bases = _bases_from_logs(ctx, logs)
    plain = Z.MSM.compute(bases, fr_arr(scalars))
    bases.precompute(c)
    xy, inf = Z.MSM.compute(bases, fr_arr(scalars))
    assert point_from_out(xy, inf) == B.msm_closed_form(logs, scalars)
    assert (xy == plain[0]).all() and inf == plain[1]
    # sub-range and batch over the same table
    off = n // 4
    sub = scalars[: n - off]
    xy, inf = Z.MSM.compute(bases, fr_arr(sub), offset=off)
    assert point_from_out(xy, inf) == B.msm_closed_form(logs[off:], sub)
    if n <= 4096:
        batch = [rand_fr(6000 + j, n) for j in range(3)]
        oxy, oinf = Z.MSM.compute_batch(bases, np.concatenate([fr_arr(b) for b in batch]), 3, n)
        for j in range(3):
            assert point_from_out(oxy[j], oinf[j]) == B.msm_closed_form(logs, batch[j])
    bases.free()


def test_msm_precomputed_special_points(ctx):
    """Table path with identity bases (flagged), repeated bases (P + P inside a bucket) and P, -P pairs."""
    n = 256
    logs = rand_fr(41, n)
    pts = [B.g1_mul(B.G1_GEN, k) for k in logs]
    pts[5] = None
    pts[100] = None
    for i in range(20, 40):
        pts[i] = pts[20]  # equal points
        logs[i] = logs[20]
    xy, inf = B.points_to_arrays(pts)
    bases = Z.Bases.upload(ctx, xy, inf).precompute(8)
    scalars = rand_fr(42, n)
    for i in range(20, 40):
        scalars[i] = scalars[20]  # same digits -> same buckets -> doubling path
    out, oinf = Z.MSM.compute(bases, fr_arr(scalars))
    exp_logs = [0 if p is None else k for p, k in zip(pts, logs)]
    assert point_from_out(out, oinf) == B.msm_closed_form(exp_logs, scalars)
    bases.free()
    logs2, sc2 = [], []
    for k, s in zip(rand_fr(43, n // 2), rand_fr(44, n // 2)):
        logs2 += [k, P - k]
        sc2 += [s, s]
    b2 = _bases_from_logs(ctx, logs2).precompute(8)
    out, oinf = Z.MSM.compute(b2, fr_arr(sc2))
    assert oinf == 1 and not out.any()
    b2.free()


TUNE_MSM_SORT = 9


@pytest.mark.parametrize("kind", ["uniform", "all_equal", "bytes_and_zeros", "offset_and_infinity"])
def test_msm_two_pass_bucket_sort(ctx, kind):
    """The two-pass shared-memory bucket sort of the window-level-table path (msm.cu 3b, from 2^18 points): the same point as
    the closed form (sum s_i k_i) G and as the atomic counting-sort scatter -- for uniform scalars, for ONE scalar repeated
    (every entry of a window in one bucket: the shared-memory cursors serialise, ranks reach the chunk size), for trace-like
    values (bytes, many zeros: most digits vanish, bins of very different sizes) and over a sub-range of bases that contains
    flagged identities."""
    from oracle import cpu_ref as R

    n = 1 << 18
    rng = np.random.default_rng(77)

    def mont_rows(count):
        a = rng.integers(0, 1 << 64, size=(count, 4), dtype=np.uint64)
        a[:, 3] &= np.uint64((1 << 60) - 1)
        return a

    k = mont_rows(n)
    if kind == "offset_and_infinity":
        k[5000:5040] = 0  # k = 0: identity bases, flagged and skipped (msm/mod.zig:407)
    bases = Z.Bases.fixed_base_mul(ctx, k)
    bases.precompute(16)  # c = 16: 2^15 buckets = 64 bins of 512, 16 windows
    if kind in ("uniform", "offset_and_infinity"):
        s = mont_rows(n)
    elif kind == "all_equal":
        s = np.repeat(mont_rows(1), n, axis=0)
    else:
        raw = np.zeros((n, 4), dtype=np.uint64)
        raw[:, 0] = rng.integers(0, 256, size=n, dtype=np.uint64)
        raw[::3, 0] = 0
        s = ctx.field_op(Z.FIELD_FR, Z.OP_TO_MONT, raw)
    off, cnt = (1 << 10, n - (1 << 10)) if kind == "offset_and_infinity" else (0, n)
    d_s = Z.DeviceBuffer.from_host(ctx, s[:cnt])
    try:
        ctx.set_tuning(TUNE_MSM_SORT, 1)
        ref = Z.MSM.compute_dev(bases, d_s.ptr, cnt, offset=off)
        ctx.set_tuning(TUNE_MSM_SORT, 0)
        got = Z.MSM.compute_dev(bases, d_s.ptr, cnt, offset=off)
        assert (got[0] == ref[0]).all() and got[1] == ref[1], "two-pass sort and atomic scatter disagree"
        dot = R.fr_dot(k[off:off + cnt], s[:cnt], threads=8)  # Montgomery limbs of sum k_i s_i
        want = B.g1_mul(B.G1_GEN, B.mont_array_to_ints(dot.reshape(1, 4), P)[0])
        assert point_from_out(*got) == want
    finally:
        ctx.set_tuning(TUNE_MSM_SORT, 0)
        d_s.free()
        bases.free()
