"""zolt_b200 -- B200-native (sm_100a) prover hot path for Zolt: BN254 G1 MSM and the sumcheck inner loop.

Layout
    csrc/        hand-written CUDA kernels + the C ABI (include/zolt_gpu.h) -> lib/libzolt_gpu.so
    host/        C++ host mirror of the reference's interfaces above the C ABI -> lib/libzolt_host.so
    _lib.py      ctypes binding of the C ABI
    api.py       thin Python mirror of the reference call sites (MSM.compute, HyperKZG.commit,
                 DensePolynomial.bindLow, sumcheck instances) used by tests/ and bench.py

There is no CPU fallback: importing works anywhere, but every operation needs the built CUDA
library and a device.
"""
from ._lib import (  # noqa: F401
    FIELD_FQ,
    FIELD_FR,
    OP_ADD,
    OP_FROM_MONT,
    OP_INV,
    OP_MUL,
    OP_NEG,
    OP_SUB,
    OP_TO_MONT,
    PAIR_ADJACENT,
    PAIR_HALF,
    SC_EQ_A_BMC,
    SC_EQ_AB_MINUS_C,
    SC_GRUEN_AB,
    SC_GRUEN_AB_MINUS_C,
    SC_STAGE4_DENSE,
    SC_EQ_DOT2,
    SC_STAGE4_GRUEN,
    SC_PROD2,
    SC_PROD3,
    SC_SUM,
    G1_WIRE_ARK_LE,
    G1_WIRE_BE,
    G1_WIRE_LE,
    TUNE_PEER_TIMEOUT_MS,
    TUNE_SC_BIND_F64,
    TUNE_SC_SMALL_PAIRS,
    TUNE_SC_TAIL_PAIRS,
    TUNE_SC_TAIL_TIMEOUT_MS,
    VEC_AFFINE_INDEX,
    VEC_AXPY,
    VEC_EQ_AB_MINUS_C,
    VEC_LIN2,
    VEC_STAGE4_TERM,
    VEC_MUL,
    ZoltGpuError,
    load_library,
)
from .api import (  # noqa: F401
    MSM,
    Bases,
    Context,
    DensePolynomial,
    DeviceBuffer,
    Dory,
    EqPolynomial,
    HyperKZG,
    MultiContext,
    PeerBox,
    SumcheckInstance,
)

__version__ = "0.1.0"
