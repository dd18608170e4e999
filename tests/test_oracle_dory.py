"""CPU tier: oracle/dory.py pinned to what the reference's own tests and comments fix (src/poly/commitment/dory.zig:1732-1750
compression round trip of the generator and of the identity; :997-1001 matrix layout examples; :740-751 file layout)."""
import struct

from oracle import bn254 as B
from oracle import dory as OD

Q = B.FQ_MOD


def _decompress(b: bytes):
    flags = b[31] & 0xC0
    if flags & 0x40:
        return None
    x = int.from_bytes(b[:31] + bytes([b[31] & 0x3F]), "little")
    y = pow((x * x * x + 3) % Q, (Q + 1) // 4, Q)  # q = 3 mod 4
    if (y <= (-y) % Q) != (flags == 0):
        y = (-y) % Q
    return (x, y)


def test_compress_roundtrip_generator_and_identity():
    assert _decompress(OD.compress_g1(B.G1_GEN)) == B.G1_GEN                     # "g1 point compression roundtrip"
    assert OD.compress_g1(None) == bytes(31) + b"\x40" and _decompress(OD.compress_g1(None)) is None  # "... identity"
    for k in (2, 3, 12345, B.FR_MOD - 1):
        p = B.g1_mul(B.G1_GEN, k)
        assert _decompress(OD.compress_g1(p)) == p
        assert _decompress(OD.compress_g1(B.g1_neg(p))) == B.g1_neg(p)
    # the sign flag really is "y > -y as standard integers"
    assert OD.compress_g1((1, 2))[31] & 0x80 == 0 and OD.compress_g1((1, Q - 2))[31] & 0x80 == 0x80


def test_matrix_layout_examples():
    # dory.zig:997-1001: num_vars=3 -> 4 cols x 2 rows, num_vars=2 -> 2 x 2, num_vars=1 -> 2 cols x 1 row;
    # :1908-1910 "Jolt with 8 coefficients uses num_columns = 4, max_num_rows = 2"
    assert OD.dory_row_split(8) == (2, 4)
    assert OD.dory_row_split(4) == (2, 2)
    assert OD.dory_row_split(2) == (1, 2)
    assert OD.dory_row_split(1) == (1, 2)
    assert OD.dory_row_split(1 << 24) == (4096, 4096)  # 4096 MSMs x 4096 points


def test_container_layouts():
    pts = [B.g1_mul(B.G1_GEN, k) for k in (1, 2, 3)]
    blob = OD.dory_srs_file(pts, 5)
    assert blob[:16] == b"JOLT_DORY_SRS_V1" and struct.unpack("<QQ", blob[16:32]) == (5, 3)
    assert int.from_bytes(blob[32:64], "little") == 1 and int.from_bytes(blob[64:96], "little") & ((1 << 254) - 1) == 2
    raw = OD.raw_binary_srs(pts)
    assert struct.unpack("<I", raw[:4])[0] == 3 and len(raw) == 4 + 3 * 64 + 128 + 64 + 128
    assert int.from_bytes(raw[4:36], "big") == 1 and int.from_bytes(raw[36:68], "big") == 2
    pt = OD.ptau_file(pts, 1)
    assert pt[:4] == b"ptau" and struct.unpack("<II", pt[4:12]) == (1, 4)
