"""GPU parity: Dory's G1 side and the SRS wire formats either side of the MSM vs oracle/dory.py.
Bit-exact: affine Montgomery limbs, infinity flags, compressed bytes."""
import numpy as np
import pytest

import zolt_b200 as Z
from oracle import bn254 as B
from oracle import dory as OD
from tests.util import P, fr_arr, rand_fr
from zolt_b200 import host as H

pytestmark = pytest.mark.gpu


def _points(seed, n, with_identity=True):
    logs = rand_fr(seed, n)
    pts = [B.g1_mul(B.G1_GEN, k) for k in logs]
    if with_identity and n > 3:
        pts[2] = None
    return logs, pts


@pytest.mark.parametrize("log_len", [0, 1, 2, 3, 7, 12, 13, 16, 20])
def test_dory_row_commitments_split_and_values(ctx, log_len):
    """DoryCommitmentScheme.commit's row MSMs (dory.zig:989-1042): sigma/nu split and every row commitment, from host and
    device tables, 1 entry up to 2^10 x 2^10."""
    n = 1 << log_len
    rows, cols = OD.dory_row_split(n)
    logs = rand_fr(600 + log_len, cols)
    g1 = Z.Bases.fixed_base_mul(ctx, fr_arr(logs))
    evals = rand_fr(700 + log_len, n)
    if n >= 64:
        evals[: n // 4] = [v % 256 for v in evals[: n // 4]]
    xy, inf, got_rows, got_cols = Z.Dory.row_commitments(g1, fr_arr(evals))
    assert (got_rows, got_cols) == ((rows, cols) if n > 1 else (1, 1))
    if log_len <= 16:
        want = OD.dory_row_commitments(logs, evals)
    else:  # 2^10 rows: check a sample against the closed form, all rows against the per-row MSM entry point
        want = None
    got = B.arrays_to_points(xy, inf)
    if want is not None:
        assert got == want
    else:
        for r in (0, 1, rows // 2, rows - 1):
            s = sum(e * k for e, k in zip(evals[r * cols:(r + 1) * cols], logs)) % P
            assert got[r] == B.g1_mul(B.G1_GEN, s)
    d = Z.DeviceBuffer.from_host(ctx, fr_arr(evals))
    xy2, inf2, r2, c2 = Z.Dory.row_commitments(g1, d.ptr, device_len=n)
    assert (xy2 == xy).all() and (inf2 == inf).all() and (r2, c2) == (got_rows, got_cols)
    d.free()
    g1.free()


def test_dory_ignores_entries_beyond_a_power_of_two(ctx):
    logs = rand_fr(611, 8)
    g1 = Z.Bases.fixed_base_mul(ctx, fr_arr(logs))
    evals = rand_fr(612, 37)  # num_vars = 5 -> 4 rows x 8 columns = the first 32 entries
    xy, inf, rows, cols = Z.Dory.row_commitments(g1, fr_arr(evals))
    assert (rows, cols) == (4, 8) and B.arrays_to_points(xy, inf) == OD.dory_row_commitments(logs, evals)
    g1.free()


def test_g1_compress(ctx):
    _, pts = _points(620, 300)
    pts[5] = (1, 2)
    pts[6] = (1, B.FQ_MOD - 2)
    xy, inf = B.points_to_arrays(pts)
    bases = Z.Bases.upload(ctx, xy, inf)
    got = bases.compress()
    assert [bytes(r) for r in got] == [OD.compress_g1(p) for p in pts]
    assert bytes(bases.compress(5, 2)[1]) == OD.compress_g1(pts[6])
    bases.free()


@pytest.mark.parametrize("fmt,wire", [("be", Z.G1_WIRE_BE), ("le", Z.G1_WIRE_LE), ("ark", Z.G1_WIRE_ARK_LE)])
def test_bases_from_wire_formats(ctx, fmt, wire):
    _, pts = _points(630, 200)
    data = b"".join(OD.g1_to_wire(p, fmt) for p in pts)
    bases = Z.Bases.from_wire(ctx, data, len(pts), wire)
    xy, inf = bases.download()
    assert B.arrays_to_points(xy, inf) == pts
    # and they are usable as MSM bases
    scalars = rand_fr(631, len(pts))
    out = Z.MSM.compute(bases, fr_arr(scalars))
    assert B.arrays_to_points(out[0].reshape(1, 8), [out[1]])[0] == B.msm(pts, scalars)
    bases.free()
    if fmt != "ark":  # the raw-binary and PTAU readers refuse points off the curve (SRSError.PointNotOnCurve)
        bad = bytearray(data)
        bad[64 * 7 + (31 if fmt == "be" else 0)] ^= 1
        with pytest.raises(Z.ZoltGpuError, match="off the curve: 7"):
            Z.Bases.from_wire(ctx, bytes(bad), len(pts), wire)


def test_srs_containers_through_the_host_mirror(ctx):
    _, pts = _points(640, 15)
    for kind, blob, nv in ((H.SRS_RAW_BINARY, OD.raw_binary_srs(pts), 0), (H.SRS_PTAU, OD.ptau_file(pts, 3), 0),
                           (H.SRS_DORY, OD.dory_srs_file(pts, 7), 7)):
        bases, err, count, max_nv, bad = H.srs_load(ctx, kind, blob)
        assert err == "Ok" and count == len(pts) and max_nv == nv and bad is None
        xy, inf = bases.download()
        assert B.arrays_to_points(xy, inf) == pts
        bases.free()
    # PTAU: at most 2^(power+1) - 1 points are taken; truncated / foreign files are refused like the reference does
    bases, err, count, _, _ = H.srs_load(ctx, H.SRS_PTAU, OD.ptau_file(pts, 2))
    assert err == "Ok" and count == 7
    bases.free()
    assert H.srs_load(ctx, H.SRS_PTAU, b"ptaX" + bytes(20))[1] == "InvalidFileFormat"
    assert H.srs_load(ctx, H.SRS_RAW_BINARY, OD.raw_binary_srs(pts)[:-1])[1] == "TruncatedData"
    assert H.srs_load(ctx, H.SRS_DORY, b"JOLT_DORY_SRS_V2" + bytes(32))[1] == "InvalidFileFormat"
    bad_blob = bytearray(OD.raw_binary_srs(pts))
    bad_blob[4 + 64 * 3 + 31] ^= 1
    assert H.srs_load(ctx, H.SRS_RAW_BINARY, bytes(bad_blob))[1:] == ("PointNotOnCurve", 15, 0, 3)


def test_dory_reduce_round_g1_ops(ctx):
    """The G1 operations of one Dory reduce round on resident vectors (dory.zig:1255, :1277-1281, :1298-1299): E1_beta over an
    SRS prefix, v1 += beta * g1_vec, and E1+ / E1- as MSMs whose BASES are the folded vector's halves (device-resident)."""
    n = 16
    g_logs, g_pts = _points(650, n, with_identity=False)
    v_logs, v_pts = _points(651, n)
    s2 = rand_fr(652, n)
    beta = 3
    g1 = Z.Bases.fixed_base_mul(ctx, fr_arr(g_logs))
    vxy, vinf = B.points_to_arrays(v_pts)
    v1 = Z.Bases.upload(ctx, vxy, vinf)
    e1_beta = Z.MSM.compute(g1, fr_arr(s2))
    assert B.arrays_to_points(e1_beta[0].reshape(1, 8), [e1_beta[1]])[0] == B.msm(g_pts, s2)
    v1.scale_add(g1, n, fr_arr([beta]))
    want_v = [B.g1_add(v, B.g1_mul(g, beta)) for v, g in zip(v_pts, g_pts)]
    xy, inf = v1.download()
    assert B.arrays_to_points(xy, inf) == want_v
    # v[i] + beta * g[i] = identity is handled: make one element cancel exactly
    w = Z.Bases.upload(ctx, *B.points_to_arrays([B.g1_neg(B.g1_mul(g_pts[0], 5)), g_pts[1]]))
    w.scale_add(g1, 2, fr_arr([5]))
    assert B.arrays_to_points(*w.download()) == [None, B.g1_add(g_pts[1], B.g1_mul(g_pts[1], 5))]
    w.free()
    # E1+ = MSM(v1_l, s2_r), E1- = MSM(v1_r, s2_l): bases = halves of the resident vector, no copy
    d_xy, d_inf = v1.dev_ptrs()
    half = n // 2
    left = Z.Bases.wrap_dev(ctx, d_xy, d_inf, half)
    right = Z.Bases.wrap_dev(ctx, d_xy + 64 * half, d_inf + half, half)
    e_plus = Z.MSM.compute(left, fr_arr(s2[half:]))
    e_minus = Z.MSM.compute(right, fr_arr(s2[:half]))
    assert B.arrays_to_points(e_plus[0].reshape(1, 8), [e_plus[1]])[0] == B.msm(want_v[:half], s2[half:])
    assert B.arrays_to_points(e_minus[0].reshape(1, 8), [e_minus[1]])[0] == B.msm(want_v[half:], s2[:half])
    for b in (left, right, v1, g1):
        b.free()
