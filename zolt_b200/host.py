"""ctypes binding of the C++ host mirror (zolt_b200/host/*.hpp -> lib/libzolt_host.so).

The host mirror is what sits above the C ABI on the CPU side: the Keccak transcript, the O(1) field
glue (UniPoly, Gruen cubic, claim updates) and the per-round loop of the stage provers
(reference src/zkvm/prover.zig:405-428).  The O(N) work is always the CUDA library.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .api import Context, _as_u64, _ptr

_host = None


def load_host() -> C.CDLL:
    global _host
    if _host is None:
        _lib.load_library()  # libzolt_host.so links against libzolt_gpu.so
        if not _lib.HOST_LIB_PATH.exists():
            raise FileNotFoundError(f"{_lib.HOST_LIB_PATH} not found: run `python -m zolt_b200.build`")
        _host = C.CDLL(str(_lib.HOST_LIB_PATH))
        _host.zolt_host_transcript_new.restype = C.c_void_p
        _host.zolt_host_transcript_new.argtypes = [C.c_char_p, C.c_size_t]
        _host.zolt_host_transcript_free.argtypes = [C.c_void_p]
        _host.zolt_host_transcript_append_bytes.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        _host.zolt_host_transcript_append_scalar.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p]
        _host.zolt_host_transcript_challenge_scalar.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p]
        _host.zolt_host_fr_op.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        _host.zolt_host_evals_to_compressed.argtypes = [C.c_void_p, C.c_void_p]
        _host.zolt_host_cubic_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _host.zolt_host_sumcheck_prove.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                                   C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        _host.zolt_host_sumcheck_prove_dev.argtypes = _host.zolt_host_sumcheck_prove.argtypes
        _host.zolt_host_sumcheck_prove_gruen.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_void_p,
                                                         C.c_void_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_void_p,
                                                         C.c_void_p, C.c_void_p]
        _host.zolt_host_sumcheck_prove_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                                           C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        _host.zolt_host_sumcheck_prove_claim.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_void_p,
                                                         C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        _host.zolt_host_sumcheck_prove_dev_claim.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_void_p,
                                                             C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        _host.zolt_host_sumcheck_prove_sharded_claim.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                                                 C.c_void_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                                                 C.c_void_p]
        _host.zolt_host_sumcheck_prove_gruen_dev.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_void_p,
                                                             C.c_void_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_void_p,
                                                             C.c_void_p, C.c_void_p]
        _host.zolt_host_sumcheck_prove_multi.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_void_p, C.c_char_p,
                                                         C.c_size_t, C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        _host.zolt_host_srs_load.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                             C.POINTER(C.c_uint64), C.POINTER(C.c_size_t)]
        _host.zolt_host_stage4_run.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _host.zolt_host_stage3_run.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)] + [C.c_void_p] * 8
        _host.zolt_host_lasso_run.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int] + [C.c_void_p] * 6
        _host.zolt_host_eq_mle.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        _host.zolt_host_horner_eval.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        _host.zolt_host_dense_evaluate.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        _host.zolt_host_batched_sumcheck.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                                     C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _host.zolt_host_hyperkzg_commit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    return _host


class Transcript:
    """Keccak Transcript(F) (reference src/transcripts/mod.zig:49-221), host side."""

    def __init__(self, domain: bytes):
        self.h = load_host().zolt_host_transcript_new(domain, len(domain))

    def append_bytes(self, data: bytes):
        load_host().zolt_host_transcript_append_bytes(self.h, data, len(data))

    def append_scalar(self, label: bytes, scalar):
        s = _as_u64(scalar, 4)
        load_host().zolt_host_transcript_append_scalar(self.h, label, len(label), _ptr(s))

    def challenge_scalar(self, label: bytes) -> np.ndarray:
        out = np.zeros(4, dtype=np.uint64)
        load_host().zolt_host_transcript_challenge_scalar(self.h, label, len(label), _ptr(out))
        return out

    def __del__(self):
        if getattr(self, "h", None) and _host is not None:
            _host.zolt_host_transcript_free(self.h)
            self.h = None


def fr_op(op: int, a, b=None) -> np.ndarray:
    a = _as_u64(a, 4)
    bb = _as_u64(b, 4) if b is not None else None
    out = np.zeros_like(a)
    load_host().zolt_host_fr_op(op, _ptr(a), _ptr(bb), _ptr(out), a.shape[0])
    return out


def evals_to_compressed(evals) -> np.ndarray:
    e = _as_u64(evals, 4)
    out = np.zeros((3, 4), dtype=np.uint64)
    load_host().zolt_host_evals_to_compressed(_ptr(e), _ptr(out))
    return out


def cubic_eval(evals, x) -> np.ndarray:
    e, xx = _as_u64(evals, 4), _as_u64(x, 4)
    out = np.zeros(4, dtype=np.uint64)
    load_host().zolt_host_cubic_eval(_ptr(e), _ptr(xx), _ptr(out))
    return out


def sumcheck_prove(ctx: Context, combiner: int, pairing: int, tables, n_extra: int = 0, domain: bytes = b"bench",
                   label: bytes = b"round", claim=None):
    """Whole dense sumcheck from HOST tables: upload, then per round GPU evals -> Keccak -> GPU fused
    bind+evals.  Returns (evals (rounds,4,4), challenges (rounds,4), final table values (k,4))."""
    tabs = [_as_u64(t, 4) for t in tables]
    n = tabs[0].shape[0]
    log_n = n.bit_length() - 1
    arr = (C.c_void_p * len(tabs))(*[t.ctypes.data for t in tabs])
    evals = np.zeros((log_n, 4, 4), dtype=np.uint64)
    chal = np.zeros((log_n, 4), dtype=np.uint64)
    final = np.zeros((len(tabs), 4), dtype=np.uint64)
    cl = _as_u64(claim, 4) if claim is not None else None
    st = load_host().zolt_host_sumcheck_prove_claim(ctx.handle, combiner, pairing, log_n, len(tabs), n_extra, arr, _ptr(cl), domain, len(domain),
                                                    label, len(label), _ptr(evals), _ptr(chal), _ptr(final))
    _lib.check(st, "zolt_host_sumcheck_prove_claim")
    return evals, chal, final


def sumcheck_prove_dev(ctx: Context, combiner: int, pairing: int, d_tables, log_n: int, n_extra: int = 0,
                       domain: bytes = b"bench", label: bytes = b"round", claim=None):
    """Same as sumcheck_prove for tables already resident on the device (list of device pointers).  `claim`: the
    instance's input claim (SumcheckInstance.input_claim); round 0 then takes p(1) from it."""
    arr = (C.c_void_p * len(d_tables))(*d_tables)
    evals = np.zeros((log_n, 4, 4), dtype=np.uint64)
    chal = np.zeros((log_n, 4), dtype=np.uint64)
    final = np.zeros((len(d_tables), 4), dtype=np.uint64)
    cl = _as_u64(claim, 4) if claim is not None else None
    st = load_host().zolt_host_sumcheck_prove_dev_claim(ctx.handle, combiner, pairing, log_n, len(d_tables), n_extra, arr, _ptr(cl), domain,
                                                        len(domain), label, len(label), _ptr(evals), _ptr(chal), _ptr(final))
    _lib.check(st, "zolt_host_sumcheck_prove_dev_claim")
    return evals, chal, final


def sumcheck_prove_gruen_dev(ctx: Context, d_tables, log_n: int, tau, claim, with_c: bool = False, domain: bytes = b"bench",
                             label: bytes = b"round"):
    """Gruen split-eq sumcheck (GruenProductProver) over device-resident tables."""
    arr = (C.c_void_p * len(d_tables))(*d_tables)
    t, cl = _as_u64(tau, 4), _as_u64(claim, 4)
    evals = np.zeros((log_n, 4, 4), dtype=np.uint64)
    chal = np.zeros((log_n, 4), dtype=np.uint64)
    final_claim = np.zeros(4, dtype=np.uint64)
    st = load_host().zolt_host_sumcheck_prove_gruen_dev(ctx.handle, int(with_c), log_n, len(d_tables), arr, _ptr(t), _ptr(cl), domain,
                                                        len(domain), label, len(label), _ptr(evals), _ptr(chal), _ptr(final_claim))
    _lib.check(st, "zolt_host_sumcheck_prove_gruen_dev")
    return evals, chal, final_claim


def stage4_run(ctx: Context, log_T: int, d_val, d_rd_wa, d_rs1_ra, d_rs2_ra, d_inc, gamma, r_cycle, phase1_rounds: int = -1,
               input_claim=None, batching_coeff=None, challenges=None, domain: bytes = b"stage4"):
    """Stage4GruenProver (src/zkvm/spartan/stage4_gruen_prover.zig) over device tables [k * T + j] (K = 128) and inc (T).
    With `challenges`: the caller-driven loop, returns (evals (rounds,4,4), challenges, finals (6,4), (input claim, last claim));
    without: `prove` against the Keccak transcript, returns the unbatched COEFFICIENTS per round and the squeezed challenges."""
    rounds = 7 + log_T
    g, rc = _as_u64(gamma, 4), _as_u64(r_cycle, 4)
    ic = None if input_claim is None else _as_u64(input_claim, 4)
    bc = None if batching_coeff is None else _as_u64(batching_coeff, 4)
    ch = None if challenges is None else _as_u64(challenges, 4)
    evals = np.zeros((rounds, 4, 4), dtype=np.uint64)
    chal = np.zeros((rounds, 4), dtype=np.uint64) if ch is None else ch.copy()
    finals = np.zeros((6, 4), dtype=np.uint64)
    claims = np.zeros((2, 4), dtype=np.uint64)
    st = load_host().zolt_host_stage4_run(ctx.handle, log_T, C.c_void_p(d_val), C.c_void_p(d_rd_wa), C.c_void_p(d_rs1_ra), C.c_void_p(d_rs2_ra),
                                          C.c_void_p(d_inc), _ptr(g), _ptr(rc), phase1_rounds, _ptr(ic) if ic is not None else None,
                                          _ptr(bc) if bc is not None else None, _ptr(ch) if ch is not None else None, domain, len(domain),
                                          _ptr(evals), _ptr(chal), _ptr(finals), _ptr(claims))
    _lib.check(st, "zolt_host_stage4_run")
    return evals, chal, finals, claims


STAGE3_COLUMNS = ("shift.unexpanded_pc", "shift.pc", "shift.is_virtual", "shift.is_first_in_sequence", "shift.is_noop",
                  "instr.left_is_rs1", "instr.rs1_value", "instr.left_is_pc", "instr.unexpanded_pc", "instr.right_is_rs2", "instr.rs2_value",
                  "instr.right_is_imm", "instr.imm", "reg.rd_write_value", "reg.rs1_value", "reg.rs2_value")


def stage3_run(ctx: Context, log_T: int, d_cols, r_outer, r_product, gammas, batching, challenges):
    """The Stage 3 provers (src/zkvm/spartan/stage3_prover.zig) over 16 device-resident trace columns (STAGE3_COLUMNS order),
    driven by the caller's challenges.  gammas: shift gamma_powers[0..4], instr gamma, reg gamma.  Returns
    (evals (rounds, 14, 4): shift 3 | instr 4 | reg 3 | combined 4, claims (6, 4): 3 input + 3 final, finals (16, 4))."""
    assert len(d_cols) == 16
    arr = (C.c_void_p * 16)(*d_cols)
    ro, rp, g, b, ch = (_as_u64(x, 4) for x in (r_outer, r_product, gammas, batching, challenges))
    evals = np.zeros((log_T, 14, 4), dtype=np.uint64)
    claims = np.zeros((6, 4), dtype=np.uint64)
    finals = np.zeros((16, 4), dtype=np.uint64)
    st = load_host().zolt_host_stage3_run(ctx.handle, log_T, arr, _ptr(ro), _ptr(rp), _ptr(g), _ptr(b), _ptr(ch), _ptr(evals), _ptr(claims),
                                          _ptr(finals))
    _lib.check(st, "zolt_host_stage3_run")
    return evals, claims, finals


def lasso_run(ctx: Context, d_indices, num_lookups: int, log_T: int, log_K: int, r_reduction, challenges=None):
    """LassoProver (src/zkvm/lasso/prover.zig) over device-resident u128 lookup indices.  With `challenges`: the caller's loop;
    without: runLassoProver with the reference's deriveChallenge.  Returns (coeffs (rounds, 3, 4), claims (rounds + 1, 4),
    challenges (rounds, 4), final_eval (4,))."""
    rounds = log_T + log_K
    rr = _as_u64(r_reduction, 4) if log_T else np.zeros((1, 4), dtype=np.uint64)
    ch = None if challenges is None else _as_u64(challenges, 4)
    coeffs = np.zeros((rounds, 3, 4), dtype=np.uint64)
    claims = np.zeros((rounds + 1, 4), dtype=np.uint64)
    chal = np.zeros((max(rounds, 1), 4), dtype=np.uint64)
    final = np.zeros(4, dtype=np.uint64)
    st = load_host().zolt_host_lasso_run(ctx.handle, C.c_void_p(d_indices), num_lookups, log_T, log_K, _ptr(rr), _ptr(ch) if ch is not None else None,
                                         _ptr(coeffs), _ptr(claims), _ptr(chal), _ptr(final))
    _lib.check(st, "zolt_host_lasso_run")
    return coeffs, claims, chal[:rounds], final


def eq_mle(r, x) -> np.ndarray:
    rr, xx = _as_u64(r, 4), _as_u64(x, 4)
    out = np.zeros(4, dtype=np.uint64)
    load_host().zolt_host_eq_mle(_ptr(rr), _ptr(xx), rr.shape[0], _ptr(out))
    return out


def horner_eval(coeffs, x) -> np.ndarray:
    cc, xx = _as_u64(coeffs, 4), _as_u64(x, 4)
    out = np.zeros(4, dtype=np.uint64)
    load_host().zolt_host_horner_eval(_ptr(cc), cc.shape[0], _ptr(xx), _ptr(out))
    return out


def dense_evaluate(ctx: Context, evals, point) -> np.ndarray:
    """DensePolynomial.evaluate (src/poly/mod.zig:73-92) through the C++ mirror (device folds)."""
    ee = _as_u64(evals, 4)
    pp = _as_u64(point, 4) if len(point) else np.empty((0, 4), dtype=np.uint64)
    out = np.zeros(4, dtype=np.uint64)
    _lib.check(load_host().zolt_host_dense_evaluate(ctx.handle, _ptr(ee), ee.shape[0], _ptr(pp), pp.shape[0], _ptr(out)), "zolt_host_dense_evaluate")
    return out


def sumcheck_prove_sharded(ctx: Context, box, combiner: int, d_tables, local_log: int, domain: bytes = b"bench",
                           label: bytes = b"round", claim=None):
    """One rank of the dense LowToHigh sumcheck over contiguous hypercube blocks (zolt_host.hpp proveSumcheckSharded):
    the round kernels exchange their sums through the NVLink mailboxes, every rank returns the whole proof."""
    world = box.world
    rounds = local_log + (world - 1).bit_length()
    arr = (C.c_void_p * len(d_tables))(*d_tables)
    evals = np.zeros((rounds, 4, 4), dtype=np.uint64)
    chal = np.zeros((rounds, 4), dtype=np.uint64)
    final = np.zeros((len(d_tables), 4), dtype=np.uint64)
    cl = _as_u64(claim, 4) if claim is not None else None
    st = load_host().zolt_host_sumcheck_prove_sharded_claim(ctx.handle, box.handle, world, combiner, local_log, len(d_tables), arr, _ptr(cl),
                                                            domain, len(domain), label, len(label), _ptr(evals), _ptr(chal), _ptr(final))
    _lib.check(st, "zolt_host_sumcheck_prove_sharded_claim")
    return evals, chal, final


def batched_sumcheck(ctx: Context, instances, claims, coeffs, challenges):
    """BatchedSumcheckProver (src/zkvm/batched_sumcheck.zig:77-430) over dense instances: `instances` is a list of
    (combiner, [host tables]); returns (compressed round polynomials (rounds, 3, 4), final claim (4,))."""
    n_inst = len(instances)
    combs = (C.c_int * n_inst)(*[c for c, _ in instances])
    tabs_flat = [_as_u64(t, 4) for _, ts in instances for t in ts]
    log_ns = (C.c_int * n_inst)(*[(_as_u64(ts[0], 4).shape[0]).bit_length() - 1 for _, ts in instances])
    n_tabs = (C.c_int * n_inst)(*[len(ts) for _, ts in instances])
    arr = (C.c_void_p * len(tabs_flat))(*[t.ctypes.data for t in tabs_flat])
    rounds = max(log_ns)
    cl, cf, ch = _as_u64(claims, 4), _as_u64(coeffs, 4), _as_u64(challenges, 4)
    comp = np.zeros((rounds, 3, 4), dtype=np.uint64)
    final = np.zeros(4, dtype=np.uint64)
    st = load_host().zolt_host_batched_sumcheck(ctx.handle, n_inst, combs, log_ns, n_tabs, arr, _ptr(cl), _ptr(cf), _ptr(ch), _ptr(comp), _ptr(final))
    _lib.check(st, "zolt_host_batched_sumcheck")
    return comp, final


def sumcheck_prove_gruen(ctx: Context, tables, tau, claim, with_c: bool = False, domain: bytes = b"bench",
                         label: bytes = b"round"):
    tabs = [_as_u64(t, 4) for t in tables]
    n = tabs[0].shape[0]
    log_n = n.bit_length() - 1
    arr = (C.c_void_p * len(tabs))(*[t.ctypes.data for t in tabs])
    t, cl = _as_u64(tau, 4), _as_u64(claim, 4)
    evals = np.zeros((log_n, 4, 4), dtype=np.uint64)
    chal = np.zeros((log_n, 4), dtype=np.uint64)
    final_claim = np.zeros(4, dtype=np.uint64)
    st = load_host().zolt_host_sumcheck_prove_gruen(ctx.handle, int(with_c), log_n, len(tabs), arr, _ptr(t), _ptr(cl), domain,
                                                    len(domain), label, len(label), _ptr(evals), _ptr(chal), _ptr(final_claim))
    _lib.check(st, "zolt_host_sumcheck_prove_gruen")
    return evals, chal, final_claim


SRS_RAW_BINARY, SRS_PTAU, SRS_DORY = 0, 1, 2
SRS_ERRORS = {0: "Ok", 1: "TruncatedData", 2: "InvalidFileFormat", 3: "UnsupportedFormat", 4: "PointNotOnCurve", 5: "Device"}


def srs_load(ctx: Context, kind: int, data: bytes):
    """srs.loadFromRawBinary / loadFromPtau / DoryCommitmentScheme.loadFromFile through the C++ mirror: the container is
    parsed on the host, the G1 array converted and checked on the device.  -> (Bases or None, error name, count,
    max_num_vars, index of the first point off the curve or None)."""
    from .api import Bases

    h = C.c_void_p()
    count, bad = C.c_size_t(), C.c_size_t()
    nv = C.c_uint64()
    err = load_host().zolt_host_srs_load(ctx.handle, kind, data, len(data), C.byref(h), C.byref(count), C.byref(nv), C.byref(bad))
    bases = Bases(ctx, h) if err == 0 and h.value else None
    return bases, SRS_ERRORS.get(err, str(err)), count.value, nv.value, (None if bad.value == (1 << 64) - 1 else bad.value)


def sumcheck_prove_multi(multi, combiner: int, tables, claim=None, domain: bytes = b"bench", label: bytes = b"round"):
    """Dense LowToHigh sumcheck of ONE process over the devices of a MultiContext (zolt_host.hpp proveSumcheckMultiDevice)."""
    tabs = [_as_u64(t, 4) for t in tables]
    log_n = tabs[0].shape[0].bit_length() - 1
    arr = (C.c_void_p * len(tabs))(*[t.ctypes.data for t in tabs])
    evals = np.zeros((log_n, 4, 4), dtype=np.uint64)
    chal = np.zeros((log_n, 4), dtype=np.uint64)
    final = np.zeros((len(tabs), 4), dtype=np.uint64)
    cl = _as_u64(claim, 4) if claim is not None else None
    st = load_host().zolt_host_sumcheck_prove_multi(multi.handle, combiner, log_n, len(tabs), arr, _ptr(cl), domain, len(domain), label,
                                                    len(label), _ptr(evals), _ptr(chal), _ptr(final))
    _lib.check(st, "zolt_host_sumcheck_prove_multi")
    return evals, chal, final
