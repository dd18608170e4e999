"""Python mirror of the reference call sites, over the C ABI.

Names follow the reference (paths relative to /root/reference):
    MSM.compute                 src/msm/mod.zig:355-372
    HyperKZG.setup / commit     src/poly/commitment/mod.zig:174-213, :239-255
    DensePolynomial.bindLow/... src/poly/mod.zig:128-175
    EqPolynomial.evals          src/poly/mod.zig:240-290
    SumcheckInstance            src/zkvm/batched_sumcheck.zig:34-74 (computeRoundPoly / bindChallenge)

Data crosses this layer as numpy uint64 arrays in the reference's own layout: field elements are
rows of 4 little-endian limbs in Montgomery form, points are rows of 8 limbs (x then y) plus a
uint8 infinity flag.  Nothing here computes on the host: every method is one or two C-ABI calls.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check


def _as_u64(a, cols: int) -> np.ndarray:
    arr = np.ascontiguousarray(a, dtype=np.uint64)
    if arr.ndim == 1 and arr.size == cols:
        arr = arr.reshape(1, cols)
    if arr.ndim != 2 or arr.shape[1] != cols:
        raise ValueError(f"expected an (n, {cols}) uint64 array, got shape {arr.shape}")
    return arr


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One device + one stream (zolt_gpu_init).  Raises if there is no GPU: no CPU path exists."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = _lib.load_library()
        h = C.c_void_p()
        check(self.lib.zolt_gpu_init(device, C.c_void_p(stream) if stream else None, C.byref(h)), "zolt_gpu_init")
        self.handle = h
        self.device = device

    def close(self):
        if getattr(self, "handle", None):
            self.lib.zolt_gpu_shutdown(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def sync(self):
        check(self.lib.zolt_gpu_sync(self.handle), "zolt_gpu_sync")

    @property
    def launch_count(self) -> int:
        out = C.c_ulonglong()
        check(self.lib.zolt_gpu_launch_count(self.handle, C.byref(out)), "zolt_gpu_launch_count")
        return out.value

    def set_tuning(self, key: int, value: int):
        check(self.lib.zolt_gpu_set_tuning(self.handle, key, value), "zolt_gpu_set_tuning")

    def timer_start(self):
        check(self.lib.zolt_timer_start(self.handle), "zolt_timer_start")

    def timer_stop_ms(self) -> float:
        ms = C.c_float()
        check(self.lib.zolt_timer_stop_ms(self.handle, C.byref(ms)), "zolt_timer_stop_ms")
        return ms.value

    def prof_enable(self, on: bool = True):
        check(self.lib.zolt_prof_enable(self.handle, int(on)), "zolt_prof_enable")

    def prof_read(self):
        """-> list of (tag, ms) recorded since the last read (see include/zolt_gpu.h for tags)."""
        tags = (C.c_int * 96)()
        ms = (C.c_float * 96)()
        n = C.c_int()
        check(self.lib.zolt_prof_read(self.handle, tags, ms, 96, C.byref(n)), "zolt_prof_read")
        return [(tags[i], ms[i]) for i in range(n.value)]

    def bench_int_pipe(self, kind: int, iters: int = 2000) -> float:
        out = C.c_double()
        check(self.lib.zolt_bench_int_pipe(self.handle, kind, iters, C.byref(out)), "zolt_bench_int_pipe")
        return out.value

    def bench_int_pipe_cfg(self, kind: int, iters: int, threads: int, blocks_per_sm: int) -> float:
        out = C.c_double()
        check(self.lib.zolt_bench_int_pipe_cfg(self.handle, kind, iters, threads, blocks_per_sm, C.byref(out)), "zolt_bench_int_pipe_cfg")
        return out.value

    # BatchOps (src/field/mod.zig:1164-1280) / parity hook for the field layer
    def field_op(self, field: int, op: int, a, b=None) -> np.ndarray:
        a = _as_u64(a, 4)
        bb = _as_u64(b, 4) if b is not None else None
        if bb is not None and bb.shape != a.shape:
            raise ValueError("operand shapes differ")
        out = np.empty_like(a)
        check(
            self.lib.zolt_field_batch_op(self.handle, field, op, _ptr(a), _ptr(bb), _ptr(out), a.shape[0]),
            "zolt_field_batch_op",
        )
        return out

    def fr_vec_dev(self, op: int, d_out: int, d_a: int = 0, d_b: int = 0, d_c: int = 0, d_e: int = 0, s=None, n: int = 0):
        """Element-wise table construction on device-resident tables (zolt_fr_vec_dev)."""
        sv = _as_u64(s, 4) if s is not None else None
        check(self.lib.zolt_fr_vec_dev(self.handle, op, C.c_void_p(d_out), C.c_void_p(d_a or None), C.c_void_p(d_b or None), C.c_void_p(d_c or None),
                                       C.c_void_p(d_e or None), _ptr(sv), n), "zolt_fr_vec_dev")

    def eq_plus_one_evals_dev(self, r, d_out: int):
        """EqPlusOnePolynomial as a table (zolt_eq_plus_one_evals): 2^len(r) elements at d_out, r big-endian."""
        rr = _as_u64(r, 4)
        check(self.lib.zolt_eq_plus_one_evals(self.handle, _ptr(rr), rr.shape[0], C.c_void_p(d_out)), "zolt_eq_plus_one_evals")

    def fr_fold_high_dev(self, d_out: int, d_in: int, d_w: int, n_lo: int, n_hi: int, scale=None):
        """out[x_lo] = scale * sum_{x_hi} in[x_lo + x_hi n_lo] w[x_hi]: the Q buffers of the prefix-suffix provers."""
        sv = _as_u64(scale, 4) if scale is not None else None
        check(self.lib.zolt_fr_fold_high_dev(self.handle, C.c_void_p(d_out), C.c_void_p(d_in), C.c_void_p(d_w), n_lo, n_hi, _ptr(sv)),
              "zolt_fr_fold_high_dev")

    def lasso_addr_round(self, d_eq_evals: int, d_indices: int, n: int, bind_bit: int, r, sum_bit: int) -> np.ndarray:
        """One LassoProver address round over device tables (zolt_lasso_addr_round) -> [sum_0, sum_1] as (2, 4) limbs."""
        rr = _as_u64(r, 4) if r is not None else None
        sums = np.zeros((2, 4), dtype=np.uint64)
        check(self.lib.zolt_lasso_addr_round(self.handle, C.c_void_p(d_eq_evals), C.c_void_p(d_indices), n, bind_bit, _ptr(rr), sum_bit,
                                             _ptr(sums)), "zolt_lasso_addr_round")
        return sums

    def poly_sum(self, d_src: int, n: int) -> np.ndarray:
        out = np.zeros(4, dtype=np.uint64)
        check(self.lib.zolt_poly_sum(self.handle, C.c_void_p(d_src), n, _ptr(out)), "zolt_poly_sum")
        return out

    def fr_inner_product_dev(self, d_a: int, d_b: int, n: int) -> np.ndarray:
        out = np.zeros(4, dtype=np.uint64)
        check(self.lib.zolt_fr_inner_product_dev(self.handle, C.c_void_p(d_a), C.c_void_p(d_b), n, _ptr(out)), "zolt_fr_inner_product_dev")
        return out

    def fr_sum_rows(self, d_rows: int, rows: int, cols: int) -> np.ndarray:
        out = np.empty((cols, 4), dtype=np.uint64)
        check(self.lib.zolt_fr_sum_rows(self.handle, C.c_void_p(d_rows), rows, cols, _ptr(out)), "zolt_fr_sum_rows")
        return out


class DeviceBuffer:
    """Raw device allocation (zolt_dev_alloc)."""

    def __init__(self, ctx: Context, nbytes: int):
        self.ctx = ctx
        self.nbytes = nbytes
        p = C.c_void_p()
        check(ctx.lib.zolt_dev_alloc(ctx.handle, nbytes, C.byref(p)), "zolt_dev_alloc")
        self.ptr = p.value

    @classmethod
    def from_host(cls, ctx: Context, arr: np.ndarray) -> "DeviceBuffer":
        arr = np.ascontiguousarray(arr)
        buf = cls(ctx, arr.nbytes)
        buf.upload(arr)
        return buf

    def upload(self, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        check(self.ctx.lib.zolt_dev_upload(self.ctx.handle, C.c_void_p(self.ptr), _ptr(arr), arr.nbytes), "zolt_dev_upload")

    def download(self, shape, dtype=np.uint64) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        check(self.ctx.lib.zolt_dev_download(self.ctx.handle, _ptr(out), C.c_void_p(self.ptr), out.nbytes), "zolt_dev_download")
        return out

    def free(self):
        if self.ptr:
            self.ctx.lib.zolt_dev_free(self.ctx.handle, C.c_void_p(self.ptr))
            self.ptr = None


class Bases:
    """Device-resident G1 bases (SetupParams.powers_of_tau_g1, commitment/mod.zig:122-140)."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self.handle = handle

    @classmethod
    def upload(cls, ctx: Context, xy, inf=None) -> "Bases":
        xy = _as_u64(xy, 8)
        infa = None if inf is None else np.ascontiguousarray(inf, dtype=np.uint8)
        h = C.c_void_p()
        check(ctx.lib.zolt_msm_bases_upload(ctx.handle, _ptr(xy), _ptr(infa), xy.shape[0], C.byref(h)), "zolt_msm_bases_upload")
        return cls(ctx, h)

    @classmethod
    def fixed_base_mul(cls, ctx: Context, scalars) -> "Bases":
        """bases[i] = scalars[i] * G with G = (1, 2)."""
        s = _as_u64(scalars, 4)
        h = C.c_void_p()
        check(ctx.lib.zolt_g1_fixed_base_mul(ctx.handle, _ptr(s), s.shape[0], C.byref(h)), "zolt_g1_fixed_base_mul")
        return cls(ctx, h)

    @classmethod
    def from_wire(cls, ctx: Context, data: bytes, n: int, fmt: int) -> "Bases":
        """n G1 points of 64 bytes straight off a file (zolt_msm_bases_upload_wire): fmt = G1_WIRE_BE (raw-binary SRS),
        G1_WIRE_LE (PTAU tauG1) or G1_WIRE_ARK_LE (Jolt Dory SRS).  Raises ZoltGpuError if a point is off the curve."""
        buf = np.frombuffer(data, dtype=np.uint8, count=64 * n) if n else np.zeros(0, dtype=np.uint8)
        h = C.c_void_p()
        bad = C.c_size_t()
        check(ctx.lib.zolt_msm_bases_upload_wire(ctx.handle, _ptr(np.ascontiguousarray(buf)), n, fmt, C.byref(h), C.byref(bad)),
              f"zolt_msm_bases_upload_wire (first point off the curve: {bad.value})")
        return cls(ctx, h)

    @classmethod
    def wrap_dev(cls, ctx: Context, d_xy: int, d_inf: int | None, n: int) -> "Bases":
        """A handle over device-resident points the caller owns (zolt_msm_bases_wrap_dev)."""
        h = C.c_void_p()
        check(ctx.lib.zolt_msm_bases_wrap_dev(ctx.handle, C.c_void_p(d_xy), C.c_void_p(d_inf or None), n, C.byref(h)), "zolt_msm_bases_wrap_dev")
        return cls(ctx, h)

    def dev_ptrs(self):
        xy, inf = C.c_void_p(), C.c_void_p()
        check(self.ctx.lib.zolt_msm_bases_dev_ptrs(self.handle, C.byref(xy), C.byref(inf)), "zolt_msm_bases_dev_ptrs")
        return xy.value, inf.value

    def compress(self, offset: int = 0, n: int | None = None) -> np.ndarray:
        """compressG1 (src/poly/commitment/dory.zig:51-76) -> (n, 32) uint8."""
        n = len(self) - offset if n is None else n
        out = np.zeros((n, 32), dtype=np.uint8)
        check(self.ctx.lib.zolt_g1_compress(self.ctx.handle, self.handle, offset, n, _ptr(out)), "zolt_g1_compress")
        return out

    def scale_add(self, g: "Bases", n: int, scalar):
        """self[i] += scalar * g[i] for i < n (Dory's v1 += beta * g1_vec, dory.zig:1277-1281)."""
        s = _as_u64(scalar, 4)
        check(self.ctx.lib.zolt_g1_scale_add(self.ctx.handle, self.handle, g.handle, n, _ptr(s)), "zolt_g1_scale_add")

    def precompute(self, window_bits: int = 0) -> "Bases":
        """zolt_msm_bases_precompute: build the window-level table (once per SRS); MSM results are unchanged."""
        check(self.ctx.lib.zolt_msm_bases_precompute(self.handle, window_bits), "zolt_msm_bases_precompute")
        return self

    def __len__(self) -> int:
        n = C.c_size_t()
        check(self.ctx.lib.zolt_msm_bases_len(self.handle, C.byref(n)), "zolt_msm_bases_len")
        return n.value

    def download(self, offset: int = 0, n: int | None = None):
        n = len(self) - offset if n is None else n
        xy = np.empty((n, 8), dtype=np.uint64)
        inf = np.empty(n, dtype=np.uint8)
        check(self.ctx.lib.zolt_msm_bases_download(self.handle, offset, n, _ptr(xy), _ptr(inf)), "zolt_msm_bases_download")
        return xy, inf

    def free(self):
        if self.handle:
            self.ctx.lib.zolt_msm_bases_free(self.handle)
            self.handle = None


class MSM:
    """msm.MSM(F, Fp) (src/msm/mod.zig:345-542)."""

    @staticmethod
    def compute(bases: Bases, scalars, offset: int = 0):
        """Sum_i scalars[i] * bases[offset + i] -> (xy (8,) uint64, infinity flag)."""
        s = _as_u64(scalars, 4) if len(scalars) else np.empty((0, 4), dtype=np.uint64)
        out = np.zeros(8, dtype=np.uint64)
        inf = C.c_uint8()
        ctx = bases.ctx
        check(ctx.lib.zolt_msm(ctx.handle, bases.handle, offset, _ptr(s), s.shape[0], _ptr(out), C.byref(inf)), "zolt_msm")
        return out, int(inf.value)

    @staticmethod
    def compute_dev(bases: Bases, d_scalars: int, n: int, offset: int = 0):
        out = np.zeros(8, dtype=np.uint64)
        inf = C.c_uint8()
        ctx = bases.ctx
        check(ctx.lib.zolt_msm_dev(ctx.handle, bases.handle, offset, C.c_void_p(d_scalars), n, _ptr(out), C.byref(inf)), "zolt_msm_dev")
        return out, int(inf.value)

    @staticmethod
    def compute_batch(bases: Bases, scalars, k: int, n: int):
        """BatchMSM.compute (src/msm/mod.zig:551-563): k scalar vectors over the same n bases."""
        s = _as_u64(np.asarray(scalars).reshape(k * n, 4), 4)
        out = np.zeros((k, 8), dtype=np.uint64)
        inf = np.zeros(k, dtype=np.uint8)
        ctx = bases.ctx
        check(ctx.lib.zolt_msm_batch(ctx.handle, bases.handle, _ptr(s), k, n, _ptr(out), _ptr(inf)), "zolt_msm_batch")
        return out, inf

    @staticmethod
    def partial_dev(bases: Bases, d_scalars: int, n: int, d_out: int, offset: int = 0):
        ctx = bases.ctx
        check(ctx.lib.zolt_msm_partial_dev(ctx.handle, bases.handle, offset, C.c_void_p(d_scalars), n, C.c_void_p(d_out)), "zolt_msm_partial_dev")

    @staticmethod
    def compute_sharded_dev(bases: Bases, d_scalars: int, n: int, box: "PeerBox", offset: int = 0):
        """ParallelMSM's point-range split across GPUs: this rank's range, partials exchanged through the NVLink
        mailboxes and summed on the device; every rank returns the same affine point."""
        out = np.zeros(8, dtype=np.uint64)
        inf = C.c_uint8()
        ctx = bases.ctx
        check(ctx.lib.zolt_msm_sharded_dev(ctx.handle, bases.handle, offset, C.c_void_p(d_scalars), n, box.handle, _ptr(out), C.byref(inf)),
              "zolt_msm_sharded_dev")
        return out, int(inf.value)

    @staticmethod
    def compute_sharded(bases: Bases, scalars, box: "PeerBox", offset: int = 0):
        """compute_sharded_dev from HOST scalars of this rank's range (upload overlapped with the multiplication)."""
        s = _as_u64(scalars, 4)
        out = np.zeros(8, dtype=np.uint64)
        inf = C.c_uint8()
        ctx = bases.ctx
        check(ctx.lib.zolt_msm_sharded(ctx.handle, bases.handle, offset, _ptr(s), s.shape[0], box.handle, _ptr(out), C.byref(inf)),
              "zolt_msm_sharded")
        return out, int(inf.value)

    @staticmethod
    def sum_partials(ctx: Context, d_partials: int, count: int):
        out = np.zeros(8, dtype=np.uint64)
        inf = C.c_uint8()
        check(ctx.lib.zolt_g1_sum_partials(ctx.handle, C.c_void_p(d_partials), count, _ptr(out), C.byref(inf)), "zolt_g1_sum_partials")
        return out, int(inf.value)


class PeerBox:
    """This rank's NVLink mailbox and the mapped mailboxes of its peers (csrc/peer.cuh).  `exchange` is any
    callable that all-gathers one 64-byte blob per rank, rank-ordered (zolt_b200.dist.connect_peer_box uses
    torch.distributed once at start-up); after that no collective library is on the data path."""

    def __init__(self, ctx: Context, rank: int, world: int, exchange):
        self.ctx, self.rank, self.world = ctx, rank, world
        h = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        st = ctx.lib.zolt_peer_box_create(ctx.handle, C.byref(h), C.cast(handle, C.c_void_p))
        self.handle = h if st == 0 else None
        # always take part in the exchange (an all-zero blob tells the peers this rank has no mailbox)
        blobs = exchange(bytes(handle) if st == 0 else bytes(64))
        check(st, "zolt_peer_box_create")
        if len(blobs) != world or any(len(b) != 64 for b in blobs):
            raise ValueError("exchange must return one 64-byte handle per rank")
        if any(b == bytes(64) for b in blobs):
            self.free()
            raise RuntimeError("a peer rank could not create its mailbox")
        allh = (C.c_ubyte * (64 * world)).from_buffer_copy(b"".join(blobs))
        check(ctx.lib.zolt_peer_box_connect(self.handle, C.cast(allh, C.c_void_p), rank, world), "zolt_peer_box_connect")

    @classmethod
    def local_group(cls, ctxs):
        """Mailboxes of several contexts of THIS process (one per rank), connected without IPC."""
        world = len(ctxs)
        boxes = []
        for rank, ctx in enumerate(ctxs):
            b = cls.__new__(cls)
            b.ctx, b.rank, b.world = ctx, rank, world
            h = C.c_void_p()
            handle = (C.c_ubyte * 64)()
            check(ctx.lib.zolt_peer_box_create(ctx.handle, C.byref(h), C.cast(handle, C.c_void_p)), "zolt_peer_box_create")
            b.handle = h
            boxes.append(b)
        arr = (C.c_void_p * world)(*[b.handle for b in boxes])
        for b in boxes:
            check(b.ctx.lib.zolt_peer_box_connect_local(b.handle, arr, b.rank, world), "zolt_peer_box_connect_local")
        return boxes

    def allgather(self, d_src: int, n_fr: int, d_out: int, transpose: bool = False):
        check(self.ctx.lib.zolt_peer_allgather(self.handle, C.c_void_p(d_src), n_fr, C.c_void_p(d_out), int(transpose)), "zolt_peer_allgather")

    def barrier(self):
        """Stream-ordered barrier across the ranks (asynchronous)."""
        check(self.ctx.lib.zolt_peer_barrier(self.handle), "zolt_peer_barrier")

    def free(self):
        if self.handle:
            check(self.ctx.lib.zolt_peer_box_free(self.handle), "zolt_peer_box_free")
            self.handle = None


class MultiContext:
    """One process, several devices (zolt_gpu_init_multi): per-rank Context / PeerBox views + the packaged sharded MSM."""

    def __init__(self, device_ids):
        self.lib = _lib.load_library()
        ids = (C.c_int * len(device_ids))(*device_ids)
        h = C.c_void_p()
        check(self.lib.zolt_gpu_init_multi(ids, len(device_ids), C.byref(h)), "zolt_gpu_init_multi")
        self.handle = h
        self.world = len(device_ids)

    def rank(self, r: int):
        """-> (Context, PeerBox) views of rank r (owned by this object: do not close / free them)."""
        ch, bh = C.c_void_p(), C.c_void_p()
        check(self.lib.zolt_multi_rank(self.handle, r, C.byref(ch), C.byref(bh)), "zolt_multi_rank")
        ctx = Context.__new__(Context)
        ctx.lib, ctx.handle, ctx.device = self.lib, ch, -1
        box = PeerBox.__new__(PeerBox)
        box.ctx, box.rank, box.world, box.handle = ctx, r, self.world, bh
        return ctx, box

    def msm(self, bases_per_rank, scalars):
        s = _as_u64(scalars, 4)
        arr = (C.c_void_p * self.world)(*[b.handle for b in bases_per_rank])
        out = np.zeros(8, dtype=np.uint64)
        inf = C.c_uint8()
        check(self.lib.zolt_multi_msm(self.handle, arr, _ptr(s), s.shape[0], _ptr(out), C.byref(inf)), "zolt_multi_msm")
        return out, int(inf.value)

    def close(self):
        if self.handle:
            self.lib.zolt_multi_shutdown(self.handle)
            self.handle = None


class Dory:
    """DoryCommitmentScheme, G1 side (src/poly/commitment/dory.zig)."""

    @staticmethod
    def row_commitments(g1: Bases, evals, device_len: int | None = None):
        """The row MSMs of commit (dory.zig:989-1042) -> (xy (rows, 8), inf (rows,), num_rows, num_cols)."""
        ctx = g1.ctx
        if device_len is None:
            e = _as_u64(evals, 4) if len(evals) else np.empty((0, 4), dtype=np.uint64)
            n, arg, fn = e.shape[0], _ptr(e), "zolt_dory_row_commitments"
        else:
            n, arg, fn = device_len, C.c_void_p(evals), "zolt_dory_row_commitments_dev"
        cap = max(1, 1 << max(0, (max(n, 1).bit_length() - 1) // 2))
        xy = np.zeros((cap, 8), dtype=np.uint64)
        inf = np.zeros(cap, dtype=np.uint8)
        rows, cols = C.c_size_t(), C.c_size_t()
        check(getattr(ctx.lib, fn)(ctx.handle, g1.handle, arg, n, _ptr(xy), _ptr(inf), C.byref(rows), C.byref(cols)), fn)
        return xy[: rows.value], inf[: rows.value], rows.value, cols.value


class HyperKZG:
    """HyperKZG(F) commit path (src/poly/commitment/mod.zig:108-255)."""

    @staticmethod
    def setup(ctx: Context, max_degree: int) -> Bases:
        h = C.c_void_p()
        check(ctx.lib.zolt_hyperkzg_setup(ctx.handle, max_degree, C.byref(h)), "zolt_hyperkzg_setup")
        return Bases(ctx, h)

    @staticmethod
    def commit(params: Bases, evals):
        """n = min(len(evals), len(powers)); MSM over the first n (commitment/mod.zig:241-249)."""
        e = _as_u64(evals, 4) if len(evals) else np.empty((0, 4), dtype=np.uint64)
        n = min(e.shape[0], len(params))
        return MSM.compute(params, e[:n])


def _hyperkzg_open(params: Bases, fn_name: str, evals_arg, n: int, point):
    ctx = params.ctx
    pt = _as_u64(point, 4) if len(point) else np.empty((0, 4), dtype=np.uint64)
    nv = pt.shape[0]
    qxy = np.zeros((max(nv, 1), 8), dtype=np.uint64)
    qinf = np.zeros(max(nv, 1), dtype=np.uint8)
    nq = C.c_int()
    final = np.zeros(4, dtype=np.uint64)
    check(getattr(ctx.lib, fn_name)(ctx.handle, params.handle, evals_arg, n, _ptr(pt), nv, _ptr(qxy), _ptr(qinf), C.byref(nq), _ptr(final)),
          fn_name)
    return qxy[: nq.value], qinf[: nq.value], final


def hyperkzg_open(params: Bases, evals, point):
    """HyperKZG.open (src/poly/commitment/mod.zig:261-324) from a host table -> (quotient xy (q,8), infinity flags (q,),
    final_eval (4,))."""
    e = _as_u64(evals, 4) if len(evals) else np.empty((0, 4), dtype=np.uint64)
    return _hyperkzg_open(params, "zolt_hyperkzg_open", _ptr(e), e.shape[0], point)


def hyperkzg_open_dev(params: Bases, d_evals: int, n: int, point):
    return _hyperkzg_open(params, "zolt_hyperkzg_open_dev", C.c_void_p(d_evals), n, point)


def hyperkzg_batch_open(params: Bases, polys, point, device_lens=None):
    """HyperKZG.batchOpen (src/poly/commitment/mod.zig:607-732).  polys: list of host tables ((n_i, 4) uint64), or device
    pointers when `device_lens` gives their lengths.  -> (evaluations (k,4), quotient xy (q,8), infinity flags (q,), final_eval
    (4,), batching challenge gamma (4,))."""
    ctx = params.ctx
    pt = _as_u64(point, 4) if len(point) else np.empty((0, 4), dtype=np.uint64)
    nv = pt.shape[0]
    k = len(polys)
    if device_lens is None:
        tabs = [_as_u64(t, 4) if len(t) else np.empty((0, 4), dtype=np.uint64) for t in polys]
        ptrs = (C.c_void_p * max(k, 1))(*[t.ctypes.data for t in tabs])
        lens = (C.c_size_t * max(k, 1))(*[t.shape[0] for t in tabs])
        fn = "zolt_hyperkzg_batch_open"
    else:
        ptrs = (C.c_void_p * max(k, 1))(*polys)
        lens = (C.c_size_t * max(k, 1))(*device_lens)
        fn = "zolt_hyperkzg_batch_open_dev"
    ev = np.zeros((max(k, 1), 4), dtype=np.uint64)
    qxy = np.zeros((max(nv, 1), 8), dtype=np.uint64)
    qinf = np.zeros(max(nv, 1), dtype=np.uint8)
    nq = C.c_int()
    final = np.zeros(4, dtype=np.uint64)
    gamma = np.zeros(4, dtype=np.uint64)
    check(getattr(ctx.lib, fn)(ctx.handle, params.handle, ptrs, lens, k, _ptr(pt), nv, _ptr(ev), _ptr(qxy), _ptr(qinf), C.byref(nq),
                               _ptr(final), _ptr(gamma)), fn)
    return ev[:k], qxy[: nq.value], qinf[: nq.value], final, gamma


HyperKZG.open = staticmethod(hyperkzg_open)
HyperKZG.batch_open = staticmethod(hyperkzg_batch_open)
HyperKZG.open_dev = staticmethod(hyperkzg_open_dev)


class DensePolynomial:
    """DensePolynomial(F) with its evaluation table resident on the device (src/poly/mod.zig:23-182)."""

    def __init__(self, ctx: Context, evaluations):
        ev = _as_u64(evaluations, 4)
        n = ev.shape[0]
        if n == 0 or n & (n - 1):
            raise ValueError("length must be a power of two")  # std.debug.assert in init (:37)
        self.ctx = ctx
        self.n = n
        self.buf = DeviceBuffer.from_host(ctx, ev)

    @property
    def num_vars(self) -> int:
        return self.n.bit_length() - 1

    def bind_low(self, r):
        """bindLow (:160-175), in place."""
        r = _as_u64(r, 4)
        check(self.ctx.lib.zolt_poly_bind_low(self.ctx.handle, C.c_void_p(self.buf.ptr), C.c_void_p(self.buf.ptr), self.n, _ptr(r)), "zolt_poly_bind_low")
        self.n //= 2

    def bind_first(self, r):
        """bindFirst (:128-149); the reference returns a new polynomial, here the table is reused."""
        r = _as_u64(r, 4)
        check(self.ctx.lib.zolt_poly_bind_first(self.ctx.handle, C.c_void_p(self.buf.ptr), C.c_void_p(self.buf.ptr), self.n, _ptr(r)), "zolt_poly_bind_first")
        self.n //= 2

    def sum(self) -> np.ndarray:
        out = np.zeros(4, dtype=np.uint64)
        check(self.ctx.lib.zolt_poly_sum(self.ctx.handle, C.c_void_p(self.buf.ptr), self.n, _ptr(out)), "zolt_poly_sum")
        return out

    def evaluations(self) -> np.ndarray:
        full = self.buf.download((self.buf.nbytes // 32, 4))
        return full[: self.n]

    def free(self):
        self.buf.free()


class EqPolynomial:
    """EqPolynomial(F).evals (src/poly/mod.zig:240-290): big-endian eq table built on the device."""

    @staticmethod
    def evals(ctx: Context, r, scale=None) -> np.ndarray:
        r = _as_u64(r, 4) if len(r) else np.empty((0, 4), dtype=np.uint64)
        n_vars = r.shape[0]
        buf = DeviceBuffer(ctx, 32 << n_vars)
        sc = _as_u64(scale, 4) if scale is not None else None
        check(ctx.lib.zolt_eq_evals(ctx.handle, _ptr(r), n_vars, _ptr(sc), C.c_void_p(buf.ptr)), "zolt_eq_evals")
        out = buf.download((1 << n_vars, 4))
        buf.free()
        return out


class SumcheckInstance:
    """k dense tables bound in lock-step on the device.

    compute_round_poly() / bind_challenge(r) are SumcheckInstance.computeRoundPolyFn /
    bindChallengeFn (src/zkvm/batched_sumcheck.zig:34-74): strictly alternating, evaluations at
    X = 0..3 (or [sum lo, sum hi] for SC_SUM, [t0, t_inf] for the Gruen combiners).
    """

    def __init__(self, ctx: Context, combiner: int, pairing: int, tables, extra_tables=(), tau=None):
        tabs = [_as_u64(t, 4) for t in list(tables) + list(extra_tables)]
        n = tabs[0].shape[0]
        if n == 0 or n & (n - 1) or any(t.shape[0] != n for t in tabs):
            raise ValueError("tables must share a power-of-two length")
        self.ctx = ctx
        self._keep = tabs
        arr = (C.c_void_p * len(tabs))(*[t.ctypes.data for t in tabs])
        h = C.c_void_p()
        check(
            ctx.lib.zolt_sc_create(ctx.handle, combiner, pairing, n.bit_length() - 1, len(tables), len(extra_tables), arr, C.byref(h)),
            "zolt_sc_create",
        )
        self.handle = h
        self._keep = None
        ne = C.c_int()
        check(ctx.lib.zolt_sc_num_evals(h, C.byref(ne)), "zolt_sc_num_evals")
        self.num_evals = ne.value
        self.num_tables = len(tabs)
        if tau is not None:
            t = _as_u64(tau, 4)
            check(ctx.lib.zolt_sc_set_split_eq(h, _ptr(t), t.shape[0]), "zolt_sc_set_split_eq")

    @classmethod
    def from_device(cls, ctx: Context, combiner: int, pairing: int, d_tables, log_n: int, n_extra: int = 0, tau=None):
        """zolt_sc_create_dev: tables are device pointers, borrowed read-only until the first bind."""
        self = cls.__new__(cls)
        self.ctx = ctx
        arr = (C.c_void_p * len(d_tables))(*d_tables)
        h = C.c_void_p()
        check(ctx.lib.zolt_sc_create_dev(ctx.handle, combiner, pairing, log_n, len(d_tables) - n_extra, n_extra, arr, C.byref(h)),
              "zolt_sc_create_dev")
        self.handle = h
        self._keep = None
        ne = C.c_int()
        check(ctx.lib.zolt_sc_num_evals(h, C.byref(ne)), "zolt_sc_num_evals")
        self.num_evals = ne.value
        self.num_tables = len(d_tables)
        if tau is not None:
            t = _as_u64(tau, 4)
            check(ctx.lib.zolt_sc_set_split_eq(h, _ptr(t), t.shape[0]), "zolt_sc_set_split_eq")
        return self

    def compute_round_poly(self) -> np.ndarray:
        out = np.empty((self.num_evals, 4), dtype=np.uint64)
        check(self.ctx.lib.zolt_sc_round_evals(self.handle, _ptr(out)), "zolt_sc_round_evals")
        return out

    def bind_challenge(self, r):
        r = _as_u64(r, 4)
        check(self.ctx.lib.zolt_sc_bind(self.handle, _ptr(r)), "zolt_sc_bind")

    def __len__(self) -> int:
        n = C.c_size_t()
        check(self.ctx.lib.zolt_sc_len(self.handle, C.byref(n)), "zolt_sc_len")
        return n.value

    def table(self, t: int) -> np.ndarray:
        out = np.empty((len(self), 4), dtype=np.uint64)
        check(self.ctx.lib.zolt_sc_download(self.handle, t, _ptr(out)), "zolt_sc_download")
        return out

    def table_dev(self, t: int) -> int:
        """Device pointer of table t's current buffer (valid until the next bind)."""
        p = C.c_void_p()
        check(self.ctx.lib.zolt_sc_table_dev(self.handle, t, C.byref(p)), "zolt_sc_table_dev")
        return p.value

    def set_claim(self, claim):
        """SumcheckInstance.input_claim: lets the first round derive p(1) = claim - p(0) too."""
        c = _as_u64(claim, 4)
        check(self.ctx.lib.zolt_sc_set_claim(self.handle, _ptr(c)), "zolt_sc_set_claim")

    def set_sharded(self, on: bool = True):
        check(self.ctx.lib.zolt_sc_set_sharded(self.handle, int(on)), "zolt_sc_set_sharded")

    def partials_dev(self) -> int:
        """Device pointer of this rank's 4 raw partial sums (128 bytes) for the all-gather."""
        p = C.c_void_p()
        check(self.ctx.lib.zolt_sc_round_partials_dev(self.handle, C.byref(p)), "zolt_sc_round_partials_dev")
        return p.value

    def finish_partials(self, d_rows: int, rows: int) -> np.ndarray:
        out = np.empty((self.num_evals, 4), dtype=np.uint64)
        check(self.ctx.lib.zolt_sc_finish_partials(self.handle, C.c_void_p(d_rows), rows, _ptr(out)), "zolt_sc_finish_partials")
        return out

    def free(self):
        if self.handle:
            self.ctx.lib.zolt_sc_free(self.handle)
            self.handle = None
