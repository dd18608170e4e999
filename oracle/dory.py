"""ORACLE (test infrastructure only): Dory's G1 side and the SRS wire formats, restated from the reference.

  compress_g1                   src/poly/commitment/dory.zig:51-76, yIsPositive :162-175
  dory_row_split / row_commitments  dory.zig:989-1042 (sigma = (num_vars + 1) / 2, nu = num_vars - sigma)
  g1 wire formats               srs.zig:65-99 (raw binary, big-endian), :616-660 (PTAU, little-endian),
                                dory.zig:828-880 (arkworks uncompressed, little-endian, flag bits cleared)
  containers                    srs.zig:256-306 (raw binary), :733-900 (PTAU), dory.zig:740-826 (Jolt Dory SRS)
Points are affine tuples (x, y) of standard integers or None for the identity (oracle/bn254.py).
"""
from __future__ import annotations

import struct

from . import bn254 as B

Q = B.FQ_MOD


def compress_g1(pt) -> bytes:
    if pt is None:
        return bytes(31) + bytes([0x40])
    x, y = pt
    out = bytearray(x.to_bytes(32, "little"))
    neg_y = (-y) % Q
    positive = y <= neg_y  # equal counts as positive
    out[31] = (out[31] & 0x3F) | (0x00 if positive else 0x80)
    return bytes(out)


def dory_row_split(length: int) -> tuple[int, int]:
    """(num_rows, num_cols) of commit()."""
    num_vars = 1 if length <= 1 else length.bit_length() - 1
    sigma = (num_vars + 1) // 2
    nu = num_vars - sigma
    return 1 << nu, 1 << sigma


def dory_row_commitments(g1_logs, evals):
    """Row MSMs of commit() over bases g1_vec[i] = g1_logs[i] * G: closed form (sum_j e[r, j] k_j) G per row."""
    if not evals:
        return []
    rows, cols = dory_row_split(len(evals))
    out = []
    for r in range(rows):
        start = r * cols
        if start >= len(evals):
            break
        row = evals[start:min(start + cols, len(evals))]
        s = sum(e * k for e, k in zip(row, g1_logs)) % B.FR_MOD
        out.append(B.g1_mul(B.G1_GEN, s))
    return out


def g1_to_wire(pt, fmt: str) -> bytes:
    """fmt: 'be' raw binary, 'le' PTAU, 'ark' Jolt Dory SRS."""
    if pt is None:
        return bytes(64)
    x, y = pt
    if fmt == "be":
        return x.to_bytes(32, "big") + y.to_bytes(32, "big")
    yb = bytearray(y.to_bytes(32, "little"))
    if fmt == "ark":
        yb[31] |= 0x80 if y > (-y) % Q else 0x00  # arkworks leaves the sign flag in place; the reader clears it
    return x.to_bytes(32, "little") + bytes(yb)


def raw_binary_srs(points) -> bytes:
    return struct.pack("<I", len(points)) + b"".join(g1_to_wire(p, "be") for p in points) + bytes(128 + 64 + 128)


def ptau_file(points, power: int, extra_sections=((7, b"contrib"),)) -> bytes:
    """A minimal snarkjs container: header (type 1), tauG1 (type 2), tauG2 (type 3, zeros) and any extra sections."""
    header = struct.pack("<I", 32) + Q.to_bytes(32, "little") + struct.pack("<II", power, power)
    secs = [(1, header)] + list(extra_sections) + [(2, b"".join(g1_to_wire(p, "le") for p in points)), (3, bytes(128 * ((1 << power) + 1)))]
    out = b"ptau" + struct.pack("<II", 1, len(secs))
    for typ, body in secs:
        out += struct.pack("<IQ", typ, len(body)) + body
    return out


def dory_srs_file(points, max_num_vars: int, g2_count: int = 2) -> bytes:
    return (b"JOLT_DORY_SRS_V1" + struct.pack("<QQ", max_num_vars, len(points)) + b"".join(g1_to_wire(p, "ark") for p in points)
            + struct.pack("<Q", g2_count) + bytes(128 * g2_count) + bytes(64 + 128))
