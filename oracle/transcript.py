"""ORACLE (test infrastructure only).  Byte-exact restatement of Zolt's Keccak `Transcript(F)`
(reference src/transcripts/mod.zig:49-221).  It is NOT standard SHA-3: bytes are XORed into
`state[position]`, the permutation runs when position reaches the rate (136), there is no padding,
and `position` is not reset by a squeeze.
"""
from __future__ import annotations

from .bn254 import FR_MOD, FR_R

KECCAK_RATE = 136  # src/transcripts/mod.zig:22
_RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000,
    0x000000000000808B, 0x0000000080000001, 0x8000000080008081, 0x8000000000008009,
    0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
    0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003,
    0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]
_ROTC = [1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44]
_PILN = [10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1]
_M = (1 << 64) - 1


def _rotl(x: int, n: int) -> int:
    return ((x << n) | (x >> (64 - n))) & _M


def keccak_f(state: bytearray) -> None:
    """keccakF (src/transcripts/mod.zig:165-219), in place on the 200-byte state."""
    st = [int.from_bytes(state[8 * i : 8 * i + 8], "little") for i in range(25)]
    for rnd in range(24):
        bc = [st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20] for i in range(5)]
        for i in range(5):
            t = bc[(i + 4) % 5] ^ _rotl(bc[(i + 1) % 5], 1)
            for j in range(i, 25, 5):
                st[j] ^= t
        t = st[1]
        for i in range(24):
            j = _PILN[i]
            tmp = st[j]
            st[j] = _rotl(t, _ROTC[i])
            t = tmp
        for row in range(0, 25, 5):
            b = st[row : row + 5]
            for i in range(5):
                st[row + i] = b[i] ^ ((~b[(i + 1) % 5]) & _M & b[(i + 2) % 5])
        st[0] ^= _RC[rnd]
    for i in range(25):
        state[8 * i : 8 * i + 8] = st[i].to_bytes(8, "little")


class Transcript:
    """Transcript(F) (src/transcripts/mod.zig:49-131).  Scalars are STANDARD integers; they are
    absorbed as their raw Montgomery limbs (appendScalar :100-109) and challenges come back as the
    standard value of fromBytes(state[0..32]) (:120-131)."""

    def __init__(self, domain: bytes):
        self.state = bytearray(200)
        self.position = 0
        self.append_bytes(domain)

    def append_bytes(self, data: bytes) -> None:
        for b in data:
            self.state[self.position] ^= b
            self.position += 1
            if self.position >= KECCAK_RATE:
                keccak_f(self.state)
                self.position = 0

    def append_scalar(self, label: bytes, scalar_std: int) -> None:
        self.append_bytes(label)
        mont = (scalar_std % FR_MOD) * FR_R % FR_MOD
        self.append_bytes(mont.to_bytes(32, "little"))

    def challenge_scalar(self, label: bytes) -> int:
        self.append_bytes(label)
        keccak_f(self.state)
        return int.from_bytes(self.state[:32], "little") % FR_MOD
