"""ctypes binding of include/zolt_gpu.h (the same symbols a `zig/zolt_gpu.zig` shim would @cImport).

The library is loaded lazily and loudly: if libzolt_gpu.so is missing or there is no CUDA device the
caller gets an exception -- there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

LIB_DIR = Path(__file__).resolve().parent / "lib"
LIB_PATH = LIB_DIR / "libzolt_gpu.so"
HOST_LIB_PATH = LIB_DIR / "libzolt_host.so"

ZOLT_OK, ZOLT_ERR_CUDA, ZOLT_ERR_ARG, ZOLT_ERR_STATE, ZOLT_ERR_NOGPU, ZOLT_ERR_PEER = 0, 1, 2, 3, 4, 5
_ERR_NAMES = {1: "ZOLT_ERR_CUDA", 2: "ZOLT_ERR_ARG", 3: "ZOLT_ERR_STATE", 4: "ZOLT_ERR_NOGPU", 5: "ZOLT_ERR_PEER"}

FIELD_FR, FIELD_FQ = 0, 1
OP_MUL, OP_ADD, OP_SUB, OP_INV, OP_FROM_MONT, OP_TO_MONT, OP_NEG = range(7)
PAIR_ADJACENT, PAIR_HALF = 0, 1
SC_SUM, SC_PROD2, SC_PROD3, SC_EQ_A_BMC, SC_EQ_AB_MINUS_C, SC_GRUEN_AB, SC_GRUEN_AB_MINUS_C, SC_STAGE4_GRUEN, SC_STAGE4_DENSE, SC_EQ_DOT2 = range(10)
TUNE_SC_SMALL_PAIRS = 1
G1_WIRE_BE, G1_WIRE_LE, G1_WIRE_ARK_LE = 0, 1, 2
TUNE_SC_BIND_F64 = 2
TUNE_PEER_TIMEOUT_MS = 3
TUNE_SC_TAIL_PAIRS = 4
TUNE_SC_TAIL_TIMEOUT_MS = 5
VEC_MUL, VEC_EQ_AB_MINUS_C, VEC_AXPY, VEC_AFFINE_INDEX, VEC_LIN2, VEC_STAGE4_TERM = range(6)


class ZoltGpuError(RuntimeError):
    def __init__(self, status: int, what: str):
        super().__init__(f"{what} failed: {_ERR_NAMES.get(status, status)}")
        self.status = status


_vp = C.c_void_p
_sz = C.c_size_t
_i = C.c_int
_u64p = C.POINTER(C.c_uint64)
_u8p = C.POINTER(C.c_uint8)

# name -> argtypes.  Every function returns int except zolt_gpu_version.
SIGNATURES = {
    "zolt_gpu_init": [_i, _vp, C.POINTER(_vp)],
    "zolt_gpu_shutdown": [_vp],
    "zolt_gpu_sync": [_vp],
    "zolt_gpu_launch_count": [_vp, C.POINTER(C.c_ulonglong)],
    "zolt_gpu_set_tuning": [_vp, _i, C.c_longlong],
    "zolt_dev_alloc": [_vp, _sz, C.POINTER(_vp)],
    "zolt_dev_free": [_vp, _vp],
    "zolt_dev_alloc_cached": [_vp, _sz, C.POINTER(C.c_void_p)],
    "zolt_dev_release_cached": [_vp, _vp, _sz],
    "zolt_dev_upload": [_vp, _vp, _vp, _sz],
    "zolt_dev_download": [_vp, _vp, _vp, _sz],
    "zolt_dev_copy": [_vp, _vp, _vp, _sz],
    "zolt_field_batch_op": [_vp, _i, _i, _vp, _vp, _vp, _sz],
    "zolt_msm_bases_upload": [_vp, _vp, _vp, _sz, C.POINTER(_vp)],
    "zolt_msm_bases_free": [_vp],
    "zolt_msm_bases_len": [_vp, C.POINTER(_sz)],
    "zolt_msm_bases_download": [_vp, _sz, _sz, _vp, _vp],
    "zolt_msm": [_vp, _vp, _sz, _vp, _sz, _vp, _vp],
    "zolt_msm_dev": [_vp, _vp, _sz, _vp, _sz, _vp, _vp],
    "zolt_msm_bases_precompute": [_vp, _i],
    "zolt_msm_batch": [_vp, _vp, _vp, _sz, _sz, _vp, _vp],
    "zolt_hyperkzg_open": [_vp, _vp, _vp, _sz, _vp, _i, _vp, _vp, C.POINTER(_i), _vp],
    "zolt_hyperkzg_open_dev": [_vp, _vp, _vp, _sz, _vp, _i, _vp, _vp, C.POINTER(_i), _vp],
    "zolt_msm_bases_upload_wire": [_vp, _vp, _sz, _i, C.POINTER(_vp), C.POINTER(_sz)],
    "zolt_msm_bases_wrap_dev": [_vp, _vp, _vp, _sz, C.POINTER(_vp)],
    "zolt_msm_bases_dev_ptrs": [_vp, C.POINTER(_vp), C.POINTER(_vp)],
    "zolt_g1_compress": [_vp, _vp, _sz, _sz, _vp],
    "zolt_g1_scale_add": [_vp, _vp, _vp, _sz, _vp],
    "zolt_dory_row_commitments": [_vp, _vp, _vp, _sz, _vp, _vp, C.POINTER(_sz), C.POINTER(_sz)],
    "zolt_dory_row_commitments_dev": [_vp, _vp, _vp, _sz, _vp, _vp, C.POINTER(_sz), C.POINTER(_sz)],
    "zolt_gpu_init_multi": [C.POINTER(_i), _i, C.POINTER(_vp)],
    "zolt_multi_shutdown": [_vp],
    "zolt_multi_size": [_vp, C.POINTER(_i)],
    "zolt_multi_rank": [_vp, _i, C.POINTER(_vp), C.POINTER(_vp)],
    "zolt_multi_msm": [_vp, C.POINTER(_vp), _vp, _sz, _vp, _vp],
    "zolt_msm_partial": [_vp, _vp, _sz, _vp, _sz, _vp],
    "zolt_hyperkzg_batch_open": [_vp, _vp, C.POINTER(_vp), C.POINTER(_sz), _sz, _vp, _i, _vp, _vp, _vp, C.POINTER(_i), _vp, _vp],
    "zolt_hyperkzg_batch_open_dev": [_vp, _vp, C.POINTER(_vp), C.POINTER(_sz), _sz, _vp, _i, _vp, _vp, _vp, C.POINTER(_i), _vp, _vp],
    "zolt_fr_inner_product_dev": [_vp, _vp, _vp, _sz, _vp],
    "zolt_fr_vec_dev": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz],
    "zolt_fr_index_map_dev": [_vp, _vp, _vp, _sz, _sz, _sz, _sz, _sz, _sz, _sz],
    "zolt_msm_partial_dev": [_vp, _vp, _sz, _vp, _sz, _vp],
    "zolt_g1_sum_partials": [_vp, _vp, _sz, _vp, _vp],
    "zolt_g1_fixed_base_mul": [_vp, _vp, _sz, C.POINTER(_vp)],
    "zolt_hyperkzg_setup": [_vp, _sz, C.POINTER(_vp)],
    "zolt_poly_bind_low": [_vp, _vp, _vp, _sz, _vp],
    "zolt_poly_bind_first": [_vp, _vp, _vp, _sz, _vp],
    "zolt_eq_evals": [_vp, _vp, _i, _vp, _vp],
    "zolt_poly_sum": [_vp, _vp, _sz, _vp],
    "zolt_eq_plus_one_evals": [_vp, _vp, _i, _vp],
    "zolt_fr_fold_high_dev": [_vp, _vp, _vp, _vp, _sz, _sz, _vp],
    "zolt_lasso_addr_round": [_vp, _vp, _vp, _sz, _i, _vp, _i, _vp],
    "zolt_sc_create": [_vp, _i, _i, _i, _i, _i, C.POINTER(_vp), C.POINTER(_vp)],
    "zolt_sc_create_dev": [_vp, _i, _i, _i, _i, _i, C.POINTER(_vp), C.POINTER(_vp)],
    "zolt_sc_set_split_eq": [_vp, _vp, _i],
    "zolt_sc_num_evals": [_vp, C.POINTER(_i)],
    "zolt_sc_round_evals": [_vp, _vp],
    "zolt_sc_bind": [_vp, _vp],
    "zolt_sc_len": [_vp, C.POINTER(_sz)],
    "zolt_sc_download": [_vp, _i, _vp],
    "zolt_sc_table_dev": [_vp, _i, C.POINTER(_vp)],
    "zolt_sc_set_claim": [_vp, _vp],
    "zolt_sc_set_sharded": [_vp, _i],
    "zolt_sc_set_peers": [_vp, _vp],
    "zolt_sc_gather_tail": [_vp, C.POINTER(_vp), C.POINTER(_sz)],
    "zolt_peer_box_create": [_vp, C.POINTER(_vp), _vp],
    "zolt_peer_box_connect": [_vp, _vp, _i, _i],
    "zolt_peer_box_connect_local": [_vp, C.POINTER(_vp), _i, _i],
    "zolt_peer_box_free": [_vp],
    "zolt_peer_allgather": [_vp, _vp, _i, _vp, _i],
    "zolt_peer_barrier": [_vp],
    "zolt_msm_sharded_dev": [_vp, _vp, _sz, _vp, _sz, _vp, _vp, _vp],
    "zolt_msm_sharded": [_vp, _vp, _sz, _vp, _sz, _vp, _vp, _vp],
    "zolt_sc_round_partials_dev": [_vp, C.POINTER(_vp)],
    "zolt_sc_finish_partials": [_vp, _vp, _sz, _vp],
    "zolt_fr_sum_rows": [_vp, _vp, _sz, _sz, _vp],
    "zolt_sc_free": [_vp],
    "zolt_bench_int_pipe": [_vp, _i, _i, C.POINTER(C.c_double)],
    "zolt_bench_int_pipe_cfg": [_vp, _i, _i, _i, _i, C.POINTER(C.c_double)],
    "zolt_prof_enable": [_vp, _i],
    "zolt_prof_read": [_vp, C.POINTER(_i), C.POINTER(C.c_float), _i, C.POINTER(_i)],
    "zolt_timer_start": [_vp],
    "zolt_timer_stop_ms": [_vp, C.POINTER(C.c_float)],
}

_lib = None


def load_library(path: os.PathLike | None = None) -> C.CDLL:
    """dlopen libzolt_gpu.so and attach prototypes.  Loading does not need a GPU (the CPU-only test
    tier checks the exported symbols); zolt_gpu_init does."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise FileNotFoundError(
            f"{p} not found: build it with `python -m zolt_b200.build` (needs nvcc). "
            "zolt_b200 has no CPU fallback."
        )
    lib = C.CDLL(str(p))
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.argtypes = argtypes
        fn.restype = _i
    lib.zolt_gpu_version.restype = C.c_char_p
    lib.zolt_gpu_version.argtypes = []
    if path is None:
        _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != ZOLT_OK:
        raise ZoltGpuError(status, what)
