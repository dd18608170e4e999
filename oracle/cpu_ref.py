"""ORACLE (test infrastructure only).  ctypes wrapper of oracle/cpu_ref.cpp."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SO = HERE / "_build" / "libzolt_cpu_ref.so"

_lib = None


def build(force: bool = False) -> Path:
    src = HERE / "cpu_ref.cpp"
    if force or not SO.exists() or SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE)] + (["-B"] if force else []), check=True, capture_output=True)
    return SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(SO))
        _lib.ref_transcript_size.restype = C.c_size_t
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u64(a, cols):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a.reshape(-1, cols)


def field_op(field: int, op: int, a, b=None) -> np.ndarray:
    a = _u64(a, 4)
    bb = _u64(b, 4) if b is not None else None
    out = np.empty_like(a)
    lib().ref_field_op(field, op, _p(a), _p(bb), _p(out), C.c_size_t(a.shape[0]))
    return out


def fr_dot(a, b, threads: int = 1) -> np.ndarray:
    """sum_i a_i * b_i over Fr (Montgomery limbs in and out)."""
    a, b = _u64(a, 4), _u64(b, 4)
    out = np.zeros(4, dtype=np.uint64)
    lib().ref_fr_dot(_p(a), _p(b), C.c_size_t(a.shape[0]), threads, _p(out))
    return out


def msm(xy, inf, scalars, threads: int = 1):
    xy = _u64(xy, 8)
    s = _u64(scalars, 4)
    infa = None if inf is None else np.ascontiguousarray(inf, dtype=np.uint8)
    out = np.zeros(8, dtype=np.uint64)
    oinf = C.c_uint8()
    lib().ref_msm(_p(xy), _p(infa), _p(s), C.c_size_t(s.shape[0]), threads, _p(out), C.byref(oinf))
    return out, int(oinf.value)


def g1_fixed_mul(scalars, threads: int = 1):
    s = _u64(scalars, 4)
    xy = np.zeros((s.shape[0], 8), dtype=np.uint64)
    inf = np.zeros(s.shape[0], dtype=np.uint8)
    lib().ref_g1_fixed_mul(_p(s), C.c_size_t(s.shape[0]), threads, _p(xy), _p(inf))
    return xy, inf


def g1_multiples(n: int):
    """bases[i] = (i+1) * G."""
    xy = np.zeros((n, 8), dtype=np.uint64)
    inf = np.zeros(n, dtype=np.uint8)
    lib().ref_g1_multiples(C.c_size_t(n), _p(xy), _p(inf))
    return xy, inf


def bind_low(table, r) -> np.ndarray:
    t = _u64(table, 4).copy()
    rr = _u64(r, 4)
    lib().ref_bind_low(_p(t), C.c_size_t(t.shape[0]), _p(rr))
    return t[: t.shape[0] // 2]


def bind_first(table, r) -> np.ndarray:
    t = _u64(table, 4)
    rr = _u64(r, 4)
    out = np.empty((t.shape[0] // 2, 4), dtype=np.uint64)
    lib().ref_bind_first(_p(t), C.c_size_t(t.shape[0]), _p(rr), _p(out))
    return out


def eq_evals(r, scale=None) -> np.ndarray:
    rr = _u64(r, 4) if len(r) else np.empty((0, 4), dtype=np.uint64)
    out = np.empty((1 << rr.shape[0], 4), dtype=np.uint64)
    sc = _u64(scale, 4) if scale is not None else None
    lib().ref_eq_evals(_p(rr), rr.shape[0], _p(sc), _p(out))
    return out


KINDS = {"sum": 0, "prod2": 1, "prod3": 2, "eq_a_bmc": 3, "eq_ab_minus_c": 4}


def sc_round(kind: str, pairing: int, tables) -> np.ndarray:
    tabs = [_u64(t, 4) for t in tables]
    arr = (C.c_void_p * len(tabs))(*[t.ctypes.data for t in tabs])
    out = np.zeros((4, 4), dtype=np.uint64)
    lib().ref_sc_round(KINDS[kind], pairing, arr, C.c_size_t(tabs[0].shape[0]), _p(out))
    return out[:2] if kind == "sum" else out


def sc_gruen_round(tables, e_out, e_in, with_c: bool) -> np.ndarray:
    tabs = [_u64(t, 4) for t in tables]
    arr = (C.c_void_p * len(tabs))(*[t.ctypes.data for t in tabs])
    eo, ei = _u64(e_out, 4), _u64(e_in, 4)
    out = np.zeros((2, 4), dtype=np.uint64)
    lib().ref_sc_gruen_round(arr, C.c_size_t(tabs[0].shape[0]), _p(eo), C.c_size_t(eo.shape[0]), _p(ei),
                             C.c_size_t(ei.shape[0]), int(with_c), _p(out))
    return out


def sc_prove(kind: str, pairing: int, tables, domain: bytes = b"bench"):
    """Whole dense sumcheck with the host Keccak transcript; returns (evals (rounds,4,4), challenges
    (rounds,4), final tables)."""
    tabs = [_u64(t, 4).copy() for t in tables]
    n = tabs[0].shape[0]
    rounds = n.bit_length() - 1
    arr = (C.c_void_p * len(tabs))(*[t.ctypes.data for t in tabs])
    evals = np.zeros((rounds, 4, 4), dtype=np.uint64)
    chal = np.zeros((rounds, 4), dtype=np.uint64)
    lib().ref_sc_prove(KINDS[kind], pairing, arr, len(tabs), C.c_size_t(n), domain, C.c_size_t(len(domain)), _p(evals), _p(chal))
    return evals, chal, [t[:1] for t in tabs]


class Transcript:
    def __init__(self, domain: bytes):
        self.buf = C.create_string_buffer(lib().ref_transcript_size())
        lib().ref_transcript_init(self.buf, domain, C.c_size_t(len(domain)))

    def append_bytes(self, data: bytes):
        lib().ref_transcript_append_bytes(self.buf, data, C.c_size_t(len(data)))

    def append_scalar(self, label: bytes, scalar_limbs):
        s = _u64(scalar_limbs, 4)
        lib().ref_transcript_append_scalar(self.buf, label, C.c_size_t(len(label)), _p(s))

    def challenge_scalar(self, label: bytes) -> np.ndarray:
        out = np.zeros(4, dtype=np.uint64)
        lib().ref_transcript_challenge_scalar(self.buf, label, C.c_size_t(len(label)), _p(out))
        return out
