#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: "BN254 MSM ms @2^24 pts (1/2/4/8 B200); sumcheck Fr elems/s & HBM GB/s".

One step = one pass of the hot path over one batch of synthetic input:
    (a) one BN254 G1 MSM of 2^24 points (the HyperKZG commit of a 2^24-entry MLE, BASELINE config 4)
    (b) one degree-3 Spartan-style sumcheck  sum eq*(Az*Bz - Cz)  over 2^24 entries with the host Keccak
        transcript in the loop (BASELINE config 3)
`value` is the MSM time in ms with inputs resident in HBM (the metric's first component, lower is
better); the sumcheck numbers ride in the "sumcheck" object of the same JSON line.  `e2e` is the same
MSM through the reference-facing C-ABI call with HOST (pinned) buffers, copies inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Multi-GPU (strong scaling, total work fixed): the MSM shards by point range and all-gathers one 128-byte
partial per rank; the sumcheck shards the hypercube in contiguous blocks and all-gathers 128 bytes of
partial evaluations per round.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "bn254_g1_msm_ms_at_2^24_points"
IMAD_PER_POINT = 42240  # SURVEY 8d convention: 16 windows x 10 modmul x 264 int32 multiply-adds
TAGS = {1: "digits", 2: "scan", 3: "scatter", 4: "accum0", 5: "edges", 6: "reduce", 7: "affine", 16: "sc_eval", 17: "sc_fused", 18: "sc_tail"}


# ------------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------------
def rand_mont(rng: np.random.Generator, n: int, out: np.ndarray | None = None) -> np.ndarray:
    """n canonical field elements as Montgomery limbs: uniform below 2^252 < r (synthetic data)."""
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    if out is not None:
        out[:] = a
        return out
    return a


def pinned_u64(n_rows: int, cols: int = 4) -> np.ndarray:
    import torch

    t = torch.empty((n_rows, cols), dtype=torch.int64, pin_memory=True)
    return t.numpy().view(np.uint64)  # the numpy view keeps the pinned tensor alive


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md "clocks line"):
    NVML polled from a thread of this process every 10 ms; `nvidia-smi -lms` only if NVML cannot be loaded."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines: list[str] = []
        self.sm: list[float] = []
        self.mx = None
        self.reasons: set[str] = set()
        self.stop_flag = threading.Event()
        self.thread = None
        self.nvml = None

    def _physical_index(self) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:  # noqa: BLE001
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        names = (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap))
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for name, bit in names:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self.stop_flag.wait(0.01)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.nvml is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.mx, "samples": len(self.sm),
                    "reasons": sorted(self.reasons), "source": "nvml, 10 ms period, timed regions only"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 100"}


def measured_peaks() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(key: str):
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(key)
        except Exception:  # noqa: BLE001
            return None
    return None


# ------------------------------------------------------------------------------------------------------
# synthetic inputs: functions of the GLOBAL index, independent of the number of GPUs
# ------------------------------------------------------------------------------------------------------
CHUNK_LOG = 16
STREAM_BASE_LOGS, STREAM_SCALARS, STREAM_EQ, STREAM_AZ, STREAM_BZ, STREAM_CZ_RANDOM, STREAM_CZ_MASK, STREAM_M20, STREAM_TRACE = range(1, 10)


def rows_global(stream: int, start: int, count: int, out: np.ndarray | None = None) -> np.ndarray:
    """Rows [start, start + count) of synthetic stream `stream` (canonical Montgomery limbs, uniform below 2^252 < r).
    Chunk c of 2^16 rows is seeded by (stream, c) alone, so every world size sees the SAME global data and a rank
    generates only its own range: the MSM and the sumcheck of the 1-, 2-, 4- and 8-GPU runs are the same problem."""
    res = out if out is not None else np.empty((count, 4), dtype=np.uint64)
    c_rows = 1 << CHUNK_LOG
    pos = 0
    c = start // c_rows
    while pos < count:
        lo = c * c_rows
        chunk = rand_mont(np.random.default_rng([0x5A17, stream, c]), c_rows)
        a = max(start, lo) - lo
        b = min(start + count, lo + c_rows) - lo
        res[pos:pos + (b - a)] = chunk[a:b]
        pos += b - a
        c += 1
    return res


def bytes_global(stream: int, start: int, count: int) -> np.ndarray:
    out = np.empty(count, dtype=np.uint8)
    c_rows = 1 << CHUNK_LOG
    pos, c = 0, start // c_rows
    while pos < count:
        lo = c * c_rows
        chunk = np.random.default_rng([0x5A17, stream, c]).integers(0, 256, size=c_rows, dtype=np.uint8)
        a = max(start, lo) - lo
        b = min(start + count, lo + c_rows) - lo
        out[pos:pos + (b - a)] = chunk[a:b]
        pos += b - a
        c += 1
    return out


def workload_config(log_n: int, sc_log_n: int) -> dict:
    """The `config` both arms print (identical in the GPU arm and the reference arm)."""
    return {
        "workload": (f"BN254 G1 MSM 2^{log_n} points (HyperKZG commit shape, BASELINE config 4) + degree-3 Spartan-style sumcheck "
                     f"sum eq*(Az*Bz-Cz) over 2^{sc_log_n} entries with the host Keccak transcript in the loop (BASELINE config 3)"),
        "inputs": ("synthetic, seeded by GLOBAL index (numpy default_rng([0x5A17, stream, chunk]), 2^16-row chunks): scalars, bases k_i*G, "
                   "tables Eq/Az/Bz uniform, Cz = Az o Bz on ~90 % of the rows; identical at every GPU count"),
    }


# ------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithms (C++ restatement, oracle/cpu_ref.cpp) on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_msm_sample(log_sample: int, threads: int, seed: int = 7):
    """Time MSM.compute / ParallelMSM.compute (reference src/msm/mod.zig:355-438, :588-653) on 2^log_sample
    points (bases generated by the same CPU code, untimed)."""
    from oracle import cpu_ref as R  # the one place bench.py may execute oracle/: the CPU baseline

    rng = np.random.default_rng(seed)
    n = 1 << log_sample
    xy, inf = R.g1_multiples(n)  # bases (i+1)*G as in src/bench.zig:251-262; scalars uniform
    scalars = rand_mont(rng, n)
    t0 = time.perf_counter()
    R.msm(xy, inf, scalars, threads=threads)
    return (time.perf_counter() - t0) * 1e3


def cpu_sumcheck_sample(log_sample: int, seed: int = 9):
    from oracle import cpu_ref as R

    rng = np.random.default_rng(seed)
    tabs = [rand_mont(rng, 1 << log_sample) for _ in range(4)]
    t0 = time.perf_counter()
    R.sc_prove("eq_ab_minus_c", 0, tabs)
    return (time.perf_counter() - t0) * 1e3


def pick_cpu_sample(threads: int, budget_s: float = 12.0, lo: int = 14, hi: int = 20) -> tuple[int, float]:
    """Largest 2^s whose MSM fits the time budget, probing upward (the algorithm is linear in n)."""
    s = lo
    ms = cpu_msm_sample(s, threads)
    while s < hi and ms * 2.2 / 1e3 < budget_s:
        s += 1
        ms = cpu_msm_sample(s, threads)
    return s, ms


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the reference is a single-process CPU program; other ranks exit 0 without work
    cores = os.cpu_count() or 1
    threads = min(8, cores)  # detectOptimalThreads caps at 8 (src/msm/mod.zig:673-678)
    log_n = args.log_n
    s, _ = pick_cpu_sample(threads, budget_s=max(4.0, 40.0 / max(1, args.steps + args.warmup)))
    s = min(s, log_n)  # a workload smaller than the sample is timed whole
    for _ in range(args.warmup):
        cpu_msm_sample(s, threads)
    t0 = time.perf_counter()
    times = [cpu_msm_sample(s, threads) for _ in range(args.steps)]
    wall = (time.perf_counter() - t0) * 1e3
    scale = float(1 << (log_n - s))
    value = statistics.mean(times) * scale
    sc_log = min(args.sc_log_n, 20)
    sc_ms = cpu_sumcheck_sample(sc_log)
    sc_scale = float(1 << (args.sc_log_n - sc_log))
    sample = (f"MSM: 2^{s} points timed with ParallelMSM.compute semantics on {threads} threads, linearly extrapolated x{scale:.0f} to 2^{log_n}; "
              f"sumcheck: 2^{sc_log} entries single-threaded (the reference has no threaded sumcheck), extrapolated x{sc_scale:.0f}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "ms", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": wall / max(1, args.steps), "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "u32x8 (256-bit Montgomery integers)", "data": "synthetic",
        "config": workload_config(log_n, args.sc_log_n),
        "cpu_baseline": {"value": value, "unit": "ms", "cores": threads, "kind": "port", "sample": sample,
                         "sumcheck_ms_extrapolated": sc_ms * sc_scale, "host_cores_available": cores},
        "e2e": {"value": value, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sumcheck": {"ms": sc_ms * sc_scale, "fr_elems_per_s": 4 * (1 << args.sc_log_n) / (sc_ms * sc_scale * 1e-3)},
        "gpu_launches": 0,
        "note": "reference cannot be compiled here (no zig toolchain): timed arm is the C++ restatement of its CPU algorithms (oracle/cpu_ref.cpp)",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# checks (untimed; the oracle is allowed here)
# ------------------------------------------------------------------------------------------------------
def _fr_ints(a) -> list[int]:
    from oracle import bn254 as B

    return B.mont_array_to_ints(np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4), B.FR_MOD)


def check_msm_closed_form(result, dot_mont_parts) -> bool:
    """result == (sum_i s_i k_i) * G for bases k_i * G; dot_mont_parts: per-rank Montgomery-form partial dot products."""
    from oracle import bn254 as B

    total = sum(_fr_ints(p)[0] for p in dot_mont_parts) % B.FR_MOD
    want = B.g1_mul(B.G1_GEN, total)
    got = B.arrays_to_points(np.asarray(result[0], dtype=np.uint64).reshape(1, 8), [int(result[1])])[0]
    return got == want


def check_sumcheck_chain(evals, chal, final, claim0: int | None, combine) -> bool:
    """The verifier's own chain: p_j(0) + p_j(1) = claim_j, claim_{j+1} = p_j(r_j), claim_final = combine(final values)."""
    from oracle import bn254 as B
    from oracle import poly as OP

    P = B.FR_MOD
    rounds = evals.shape[0]
    e0 = _fr_ints(evals[0])
    claim = (e0[0] + e0[1]) % P
    if claim0 is not None and claim != claim0 % P:
        return False
    for j in range(rounds):
        e = _fr_ints(evals[j])
        if (e[0] + e[1]) % P != claim:
            return False
        claim = OP.evaluate_cubic_at_point(e, _fr_ints(chal[j])[0])
    f = [_fr_ints(final[t])[0] for t in range(final.shape[0])]
    return claim == combine(f) % P


# ------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------
def run_b200(args):
    import hashlib

    import torch
    import torch.distributed as dist

    import zolt_b200 as Z
    from zolt_b200 import host as H

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"  # NCCL's version banner goes to stdout, which must carry the JSON line only
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    ctx = Z.Context(local_rank, stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_rows(row: np.ndarray) -> list[np.ndarray]:
        """one (4,) uint64 row per rank, rank-ordered (setup/check plumbing, never inside a timed region)"""
        if world == 1:
            return [np.asarray(row, dtype=np.uint64)]
        t = torch.from_numpy(np.asarray(row, dtype=np.uint64).view(np.int64).copy()).cuda()
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [o.cpu().numpy().view(np.uint64) for o in out]

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        return e

    def time_ms(fn, reps: int, warm: int = 1) -> float:
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(reps):
            a = ev()
            fn()
            b = ev()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.mean(ts)

    log_n, sc_log_n = args.log_n, args.sc_log_n
    shift = (world - 1).bit_length()
    if (1 << shift) != world:
        raise SystemExit("world size must be a power of two")
    n_local = 1 << (log_n - shift)
    sc_local_log = sc_log_n - shift
    sc_local = 1 << sc_local_log
    K_TABLES = 4
    pt0, sc0 = rank * n_local, rank * sc_local  # this rank's global ranges (MSM: point range; sumcheck: hypercube block)

    # ---- setup (untimed): bases k_i*G generated on the device, host inputs in pinned memory -------------
    h_logs = rows_global(STREAM_BASE_LOGS, pt0, n_local)
    bases = Z.Bases.fixed_base_mul(ctx, h_logs)
    h_scalars = rows_global(STREAM_SCALARS, pt0, n_local, pinned_u64(n_local))
    d_scalars = Z.DeviceBuffer.from_host(ctx, h_scalars)
    extras = {}
    # plain bases first (MSM.compute(bases, scalars) on arbitrary bases), then the per-SRS table
    if not args.no_table:
        extras["msm_no_table_ms"] = time_ms(lambda: Z.MSM.compute_dev(bases, d_scalars.ptr, n_local), 3)
        ctx.sync()
        t0 = time.perf_counter()
        bases.precompute(0)  # once per SRS (long-lived SetupParams): window-level table, 13 additions per point
        ctx.sync()
        extras["precompute_ms"] = (time.perf_counter() - t0) * 1e3
    msm_windows = 16 if args.no_table or n_local < (1 << 15) else (13 if n_local >= (1 << 20) else 16)

    h_tabs = [rows_global(s, sc0, sc_local, pinned_u64(sc_local)) for s in (STREAM_EQ, STREAM_AZ, STREAM_BZ)]
    d_tabs = [Z.DeviceBuffer.from_host(ctx, t) for t in h_tabs]
    # Cz = Az o Bz on ~90 % of the rows (a mostly satisfied R1CS), uniform elsewhere: the product runs on the device (input
    # generation, untimed), the patch on the host
    d_cz = Z.DeviceBuffer(ctx, sc_local * 32)
    ctx.fr_vec_dev(Z.VEC_MUL, d_cz.ptr, d_tabs[1].ptr, d_tabs[2].ptr, n=sc_local)
    h_cz = pinned_u64(sc_local)
    h_cz[:] = d_cz.download((sc_local, 4))
    bad = bytes_global(STREAM_CZ_MASK, sc0, sc_local) < 26
    h_cz[bad] = rows_global(STREAM_CZ_RANDOM, sc0, sc_local)[bad]
    d_cz.upload(h_cz)
    h_tabs.append(h_cz)
    d_tabs.append(d_cz)
    d_tab_ptrs = [b.ptr for b in d_tabs]
    # the instance's input claim sum eq*(az*bz - cz) (SumcheckInstance.input_claim: the protocol hands it to the prover)
    d_tmp = Z.DeviceBuffer(ctx, sc_local * 32)
    ctx.fr_vec_dev(Z.VEC_EQ_AB_MINUS_C, d_tmp.ptr, d_tabs[1].ptr, d_tabs[2].ptr, d_tabs[3].ptr, d_tabs[0].ptr, n=sc_local)
    part_sum = np.zeros(4, dtype=np.uint64)
    Z._lib.check(ctx.lib.zolt_poly_sum(ctx.handle, __import__("ctypes").c_void_p(d_tmp.ptr), sc_local, part_sum.ctypes.data_as(__import__("ctypes").c_void_p)), "zolt_poly_sum")
    d_tmp.free()
    claim = np.zeros(4, dtype=np.uint64)
    for row in gather_rows(part_sum):
        claim = H.fr_op(1, claim.reshape(1, 4), row.reshape(1, 4))[0]

    part = torch.zeros(128, dtype=torch.uint8, device="cuda")
    gathered = torch.zeros(128 * world, dtype=torch.uint8, device="cuda")
    ev_part = torch.zeros(128, dtype=torch.uint8, device="cuda")
    ev_gath = torch.zeros(128 * world, dtype=torch.uint8, device="cuda")

    # multi-GPU exchange: NVLink peer mailboxes written by the kernels themselves (default), or NCCL all-gathers
    # driven from the host (--nccl-exchange, the baseline the mailbox replaces)
    box = None
    if world > 1 and not args.nccl_exchange:
        from zolt_b200 import dist as ZD

        try:
            box = ZD.connect_peer_box(ctx, rank, world)  # raises on every rank alike if any rank cannot map its peers
        except RuntimeError as e:  # e.g. CUDA IPC unavailable in this container: fall back to the NCCL exchange, visibly
            box = None
            if rank == 0:
                print(f"[bench] {e}; using the NCCL exchange", file=sys.stderr, flush=True)

    def msm_resident():
        if world == 1:
            return Z.MSM.compute_dev(bases, d_scalars.ptr, n_local)
        if box is not None:
            return Z.MSM.compute_sharded_dev(bases, d_scalars.ptr, n_local, box)
        with torch.cuda.stream(stream):
            Z.MSM.partial_dev(bases, d_scalars.ptr, n_local, part.data_ptr())
            dist.all_gather_into_tensor(gathered, part)
            return Z.MSM.sum_partials(ctx, gathered.data_ptr(), world)

    def sumcheck_resident():
        if world == 1:
            return H.sumcheck_prove_dev(ctx, Z.SC_EQ_AB_MINUS_C, Z.PAIR_ADJACENT, d_tab_ptrs, sc_local_log, claim=claim)
        if box is not None:
            return H.sumcheck_prove_sharded(ctx, box, Z.SC_EQ_AB_MINUS_C, d_tab_ptrs, sc_local_log, claim=claim)
        from zolt_b200 import dist as ZD

        return ZD.sharded_sumcheck_gpu(ctx, stream, Z.SC_EQ_AB_MINUS_C, d_tab_ptrs, sc_local_log, world, ev_part, ev_gath)

    def msm_e2e(scalars=h_scalars):
        if world == 1:
            return Z.MSM.compute(bases, scalars)
        if box is not None:
            return Z.MSM.compute_sharded(bases, scalars, box)
        with torch.cuda.stream(stream):
            d_scalars.upload(scalars)
            Z.MSM.partial_dev(bases, d_scalars.ptr, n_local, part.data_ptr())
            dist.all_gather_into_tensor(gathered, part)
            return Z.MSM.sum_partials(ctx, gathered.data_ptr(), world)

    def sumcheck_e2e(tabs=h_tabs):
        if world == 1:
            return H.sumcheck_prove(ctx, Z.SC_EQ_AB_MINUS_C, Z.PAIR_ADJACENT, tabs, claim=claim)
        for b, t in zip(d_tabs, tabs):
            b.upload(t)
        return sumcheck_resident()

    # ---- warm-up -----------------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 1)):
        r_msm = msm_resident()
        r_sc = sumcheck_resident()
    if args.warmup:
        msm_e2e()
        sumcheck_e2e()
    int_peak = ctx.bench_int_pipe(0, 4000) if rank == 0 else 0.0  # IMAD/s, measured live (untimed region)

    # ---- timed loop 1: inputs resident in HBM --------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count
    ctx.prof_enable(True)
    msm_ms, sc_ms, phase_log = [], [], []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        if box is not None:
            box.barrier()  # align the ranks' streams: time the MSM, not the arrival skew of the previous phase
        a = ev()
        r_msm = msm_resident()
        b = ev()
        r_sc = sumcheck_resident()
        c = ev()
        c.synchronize()
        msm_ms.append(a.elapsed_time(b))
        sc_ms.append(b.elapsed_time(c))
        phase_log.append(ctx.prof_read())
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    ctx.prof_enable(False)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed loop 2: end to end through the C ABI with host buffers (pinned) ----------------------------------
    barrier()
    e2e_msm, e2e_sc = [], []
    for _ in range(args.steps):
        if box is not None:
            box.barrier()
        a = ev()
        msm_e2e()
        b = ev()
        sumcheck_e2e()
        c = ev()
        c.synchronize()
        e2e_msm.append(a.elapsed_time(b))
        e2e_sc.append(b.elapsed_time(c))
    barrier()
    # ... and once more from PAGEABLE host memory, which is what a Zig caller's []const F is (2 steps)
    pg_scalars = np.array(h_scalars, copy=True)
    pg_tabs = [np.array(t, copy=True) for t in h_tabs]
    pg_msm, pg_sc = [], []
    for _ in range(2):
        if box is not None:
            box.barrier()
        a = ev()
        msm_e2e(pg_scalars)
        b = ev()
        sumcheck_e2e(pg_tabs)
        c = ev()
        c.synchronize()
        pg_msm.append(a.elapsed_time(b))
        pg_sc.append(b.elapsed_time(c))
    del pg_tabs
    barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    v_msm = max_over_ranks(statistics.mean(msm_ms))
    v_sc = max_over_ranks(statistics.mean(sc_ms))
    v_wall = max_over_ranks(wall_ms)
    v_e2e_msm = max_over_ranks(statistics.mean(e2e_msm))
    v_e2e_sc = max_over_ranks(statistics.mean(e2e_sc))
    v_pg_msm = max_over_ranks(statistics.mean(pg_msm))
    v_pg_sc = max_over_ranks(statistics.mean(pg_sc))

    # ---- checks (untimed): every rank contributes its part of the closed form, rank 0 decides ---------------------
    from oracle import cpu_ref as R  # checker only

    cores = os.cpu_count() or 1
    dot_parts = gather_rows(R.fr_dot(h_logs, h_scalars, threads=max(1, min(8, cores // max(1, world)))))
    sc_evals, sc_chal, sc_final = r_sc
    digest_src = hashlib.sha256()
    for arr in (np.asarray(r_msm[0], dtype=np.uint64), np.asarray([r_msm[1]], dtype=np.uint64), sc_evals, sc_chal, sc_final):
        digest_src.update(np.ascontiguousarray(arr).tobytes())
    digests = [digest_src.hexdigest()]
    if world > 1:
        all_d = [None] * world
        dist.all_gather_object(all_d, digests[0])
        digests = all_d

    if rank == 0:
        checks = {
            "msm_equals_closed_form": check_msm_closed_form(r_msm, dot_parts),
            "sumcheck_verifier_chain": check_sumcheck_chain(sc_evals, sc_chal, sc_final, _fr_ints(claim)[0], lambda f: f[0] * (f[1] * f[2] - f[3])),
            "all_ranks_same_digest": len(set(digests)) == 1,
            "digest": digests[0],
            "digest_of": "sha256(MSM affine x|y limbs, infinity flag, all round evaluations, all challenges, final bound values); identical at every GPU count",
        }
        # per-kernel device times (CUDA events inside the library, on the launching stream)
        acc = {}
        for log in phase_log:
            for tag, ms in log:
                acc.setdefault(TAGS.get(tag, str(tag)), []).append(ms)
        steps = max(1, len(phase_log))
        phase_ms = {k: sum(v) / steps for k, v in acc.items()}
        accum_ms = phase_ms.get("accum0", float("nan"))
        sc_kernel_ms = phase_ms.get("sc_eval", 0.0) + phase_ms.get("sc_fused", 0.0) + phase_ms.get("sc_tail", 0.0)
        hbm_peak, peak_src = measured_peaks()
        achieved_imad = IMAD_PER_POINT * n_local / (accum_ms * 1e-3) / 1e12
        # windows actually accumulated x multiply-adds of one mixed addition: 8 products x 128 + 2 squarings x 72 + 9 reductions
        # x 136 (y3 shares one reduction between its two products)
        executed_imad = msm_windows * 2392 * n_local / (accum_ms * 1e-3) / 1e12
        sc_bytes = 128 * K_TABLES * sc_local
        sc_gbs = sc_bytes / (sc_kernel_ms * 1e-3) / 1e9 if sc_kernel_ms else float("nan")
        n_sc = 1 << sc_log_n
        cfg = workload_config(log_n, sc_log_n)
        cfg.update({
            "sharding": ("MSM by point range, sumcheck by contiguous hypercube blocks; exchange: "
                         + ("NVLink peer mailboxes written by the kernels (no collective library on the data path)" if box is not None
                            else "NCCL all-gather per round")) if world > 1 else "single GPU",
            "l2": "inputs (1.5 GiB MSM, 2 GiB sumcheck) exceed the 126 MB L2; no flush needed",
            "msm_bases": "window-level table built once per SRS in setup (zolt_msm_bases_precompute, c=20, 13 windows); plain-bases time and table build time in `extras`" if not args.no_table else "plain bases (16 windows)",
            "sumcheck_claim": "input claim supplied (SumcheckInstance.input_claim): round 0 takes p(1) from p(0) + p(1) = claim",
        })
        line = {
            "metric": METRIC, "value": v_msm, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": v_wall / args.steps, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32x8 (256-bit Montgomery integers)", "data": "synthetic",
            "config": cfg,
            "roofline": {
                "kernel": "msm_accum_points_kernel (bucket accumulation, mixed additions)", "bound": "int32",
                "achieved": achieved_imad, "peak": int_peak / 1e12, "unit": "TIMAD/s", "frac": achieved_imad / (int_peak / 1e12) if int_peak else None,
                "peak_source": "IMAD micro-benchmark run live in this process (zolt_bench_int_pipe kind 0)",
                "work_per_launch": f"{IMAD_PER_POINT} IMAD/point x {n_local} points (SURVEY 8d convention: 16 windows x 10 modmul x 264)",
                "achieved_executed": executed_imad, "frac_executed": executed_imad / (int_peak / 1e12) if int_peak else None,
                "note": f"the kernel accumulates {msm_windows} windows per point, not the convention's 16, and spends 2392 multiply-adds per mixed addition (dedicated squarings, one shared reduction), not 2640: frac > 1 means fewer operations than the convention's algorithm, frac_executed is the pipe utilisation of the operations actually issued",
                "kernel_ms": accum_ms, "traffic": ncu_traffic("msm_accum_points_kernel") if world == 1 else None,
            },
            "roofline_sumcheck": {
                "kernel": "sc_round_kernel + sc_tail_kernel (round-0 eval, fused bind/eval rounds, resident tail: all launches of one sumcheck)", "bound": "hbm",
                "achieved": sc_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": sc_gbs / hbm_peak, "peak_source": peak_src,
                "frac_of_8000_spec": sc_gbs / 8000.0, "bytes_per_sumcheck": sc_bytes, "kernel_ms": sc_kernel_ms,
                "traffic": ncu_traffic("sc_round_kernel") if world == 1 else None,
                "binding_unit": "integer multiplier (ncu sm__pipe_fmaheavy_cycles_active 79-83 % on the fused round, DRAM 50 %, traffic = "
                                "algorithmic bytes: profiles/r2_v23_sc_round_full.txt); at 100 % of that pipe the kernels would reach 0.55 of "
                                "the HBM peak",
            },
            "sumcheck": {
                "ms": v_sc, "kernel_ms": sc_kernel_ms, "fr_elems_per_s": K_TABLES * n_sc / (v_sc * 1e-3), "hbm_gbs_wall": 128 * K_TABLES * n_sc / (v_sc * 1e-3) / 1e9,
                "e2e_ms": v_e2e_sc, "e2e_fr_elems_per_s": K_TABLES * n_sc / (v_e2e_sc * 1e-3),
            },
            "msm_phases_ms": {k: round(v, 4) for k, v in phase_ms.items() if not k.startswith("sc_")},
            "sc_round_kernel_ms": [round(ms, 4) for tag, ms in phase_log[-1] if tag in (16, 17, 18)] if phase_log else [],
            "sc_round_kernel_note": "one entry per launch: round 0, the fused rounds, and last the resident tail kernel (all rounds of at most 2^12 pairs, host round trips included)",
            "e2e": {"value": v_e2e_msm, "unit": "ms", "h2d_bytes_per_step": int(n_local * 32 + K_TABLES * sc_local * 32),
                    "d2h_bytes_per_step": int(65 + sc_log_n * (128 + 32) + K_TABLES * 32), "sumcheck_ms": v_e2e_sc,
                    "host_buffers": "pinned", "pageable_ms": v_pg_msm, "pageable_sumcheck_ms": v_pg_sc},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "checks": checks,
            "extras": {k: round(v, 4) for k, v in extras.items()},
        }
        if not all(v for k, v in checks.items() if isinstance(v, bool)):
            line["error"] = "a result check failed: see `checks`"
    if world == 1 and not args.no_extras:
        try:
            line["extras"].update(extra_legs(ctx, Z, H, R, bases, h_logs, h_scalars, d_tab_ptrs, int_peak, time_ms, args))
        except Exception as e:  # noqa: BLE001 - an extra leg must never cost the headline line; the failure is reported in it
            line["extras"]["error"] = f"{type(e).__name__}: {e}"
    if rank == 0:
        if world == 1 and not args.no_cpu:
            threads = min(8, cores)
            s, ms = pick_cpu_sample(threads, budget_s=10.0)
            if s > log_n:  # a workload smaller than the sample is timed whole
                s = log_n
                ms = cpu_msm_sample(s, threads)
            scale = float(1 << (log_n - s))
            sc_s = min(sc_log_n, 20)
            sc_cpu = cpu_sumcheck_sample(sc_s)
            line["cpu_baseline"] = {
                "value": ms * scale, "unit": "ms", "cores": threads, "kind": "port",
                "sample": f"MSM 2^{s} points on {threads} threads (ParallelMSM.compute semantics, src/msm/mod.zig:588-653), {ms:.0f} ms, linear x{scale:.0f} to 2^{log_n}; reference cannot be built here (no zig)",
                "sumcheck_ms": sc_cpu * (1 << (sc_log_n - sc_s)), "sumcheck_sample": f"2^{sc_s} entries single-threaded, {sc_cpu:.0f} ms, linear x{1 << (sc_log_n - sc_s)}",
                "host_cores_available": cores,
            }
        print(json.dumps(line), flush=True)
    if box is not None:
        barrier()
        box.free()
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


def extra_legs(ctx, Z, H, R, bases, h_logs, h_scalars, d_tab_ptrs, int_peak, time_ms, args) -> dict:
    """Single-GPU only: the other BASELINE configs and workload variants, each timed on the device (CUDA events on the
    launching stream, 1 warm-up + 3-5 runs) and checked: MSM 2^16 (config 1), the trace-like scalar distribution of config 4,
    the degree-2 Eq*MLE sumcheck over 2^20 entries (config 2, the latency-bound case) and the Gruen form of config 3."""
    from oracle import bn254 as B

    out = {}
    hbm_peak, _ = measured_peaks()
    n = h_scalars.shape[0]
    # config 1: MSM 2^16 over the first 2^16 bases
    n16 = min(1 << 16, n)
    d16 = Z.DeviceBuffer.from_host(ctx, h_scalars[:n16])
    res = [None]

    def run16():
        res[0] = Z.MSM.compute_dev(bases, d16.ptr, n16)

    ms = time_ms(run16, 5)
    out["msm_2^16_ms"] = ms
    out["msm_2^16_frac_int32_convention"] = IMAD_PER_POINT * n16 / (ms * 1e-3) / int_peak if int_peak else None
    out["msm_2^16_ok"] = check_msm_closed_form(res[0], [R.fr_dot(h_logs[:n16], h_scalars[:n16])])
    d16.free()
    # config 4, trace-like scalars: half of the entries byte-sized (a third of those zero), a quarter 64-bit, a quarter zero
    rng = np.random.default_rng([0x5A17, STREAM_TRACE])
    small = rng.integers(0, 256, size=n, dtype=np.uint64)
    small[::3] = 0
    kind = rng.integers(0, 4, size=n)
    vals = np.where(kind < 2, small, np.where(kind == 2, rng.integers(0, 1 << 63, size=n, dtype=np.uint64), 0)).astype(np.uint64)
    raw = np.zeros((n, 4), dtype=np.uint64)
    raw[:, 0] = vals
    step = 1 << 20
    trace = np.empty_like(raw)
    for lo in range(0, n, step):  # into Montgomery form on the device (zolt_field_batch_op TO_MONT): input generation
        trace[lo:lo + step] = ctx.field_op(Z.FIELD_FR, Z.OP_TO_MONT, raw[lo:lo + step])
    d_tr = Z.DeviceBuffer.from_host(ctx, trace)

    def run_tr():
        res[0] = Z.MSM.compute_dev(bases, d_tr.ptr, n)

    out["msm_trace_like_ms"] = time_ms(run_tr, 3)
    out["msm_trace_like_ok"] = check_msm_closed_form(res[0], [R.fr_dot(h_logs, trace, threads=min(8, os.cpu_count() or 1))])
    d_tr.free()
    # config 2: degree-2 Eq*M over 2^20 entries, host Keccak in the loop (latency-bound: report kernels and wall)
    l2 = min(20, args.sc_log_n)
    tau = rows_global(STREAM_EQ, 0, l2)
    d_eq = Z.DeviceBuffer(ctx, 32 << l2)
    import ctypes as C

    Z._lib.check(ctx.lib.zolt_eq_evals(ctx.handle, tau.ctypes.data_as(C.c_void_p), l2, None, C.c_void_p(d_eq.ptr)), "zolt_eq_evals")
    d_m = Z.DeviceBuffer.from_host(ctx, rows_global(STREAM_M20, 0, 1 << l2))

    def run_c2():
        res[0] = H.sumcheck_prove_dev(ctx, Z.SC_PROD2, Z.PAIR_ADJACENT, [d_eq.ptr, d_m.ptr], l2)

    run_c2()
    ctx.prof_enable(True)
    wall = time_ms(run_c2, 5, warm=0)
    kms = sum(ms for tag, ms in ctx.prof_read() if tag in (16, 17, 18)) / 5
    ctx.prof_enable(False)
    out[f"sumcheck_deg2_2^{l2}_ms"] = wall
    out[f"sumcheck_deg2_2^{l2}_kernel_ms"] = kms
    out[f"sumcheck_deg2_2^{l2}_frac_hbm_kernels"] = 128 * 2 * (1 << l2) / (kms * 1e-3) / 1e9 / hbm_peak if kms else None
    out[f"sumcheck_deg2_2^{l2}_ok"] = check_sumcheck_chain(*res[0], None, lambda f: f[0] * f[1])
    d_eq.free()
    d_m.free()
    # config 3, Gruen form: eq(tau, .) never materialised (GRUEN_AB_MINUS_C over Az, Bz, Cz + split-eq weights)
    lg = args.sc_log_n
    tau_g = rows_global(STREAM_EQ, 1 << 16, lg)
    d_eqg = Z.DeviceBuffer(ctx, 32 << lg)  # only to compute the input claim sum eq(tau, x)(az bz - cz)(x), untimed
    Z._lib.check(ctx.lib.zolt_eq_evals(ctx.handle, tau_g.ctypes.data_as(C.c_void_p), lg, None, C.c_void_p(d_eqg.ptr)), "zolt_eq_evals")
    ctx.fr_vec_dev(Z.VEC_EQ_AB_MINUS_C, d_eqg.ptr, d_tab_ptrs[1], d_tab_ptrs[2], d_tab_ptrs[3], d_eqg.ptr, n=1 << lg)
    claim_g = np.zeros(4, dtype=np.uint64)
    Z._lib.check(ctx.lib.zolt_poly_sum(ctx.handle, C.c_void_p(d_eqg.ptr), 1 << lg, claim_g.ctypes.data_as(C.c_void_p)), "zolt_poly_sum")
    d_eqg.free()

    def run_g():
        res[0] = H.sumcheck_prove_gruen_dev(ctx, d_tab_ptrs[1:4], lg, tau_g, claim_g, with_c=True)

    run_g()
    ctx.prof_enable(True)
    wall = time_ms(run_g, 3, warm=0)
    kms = sum(ms for tag, ms in ctx.prof_read() if tag in (16, 17, 18)) / 3
    ctx.prof_enable(False)
    out[f"sumcheck_gruen_2^{lg}_ms"] = wall
    out[f"sumcheck_gruen_2^{lg}_kernel_ms"] = kms
    out[f"sumcheck_gruen_2^{lg}_frac_hbm_kernels"] = 128 * 3 * (1 << lg) / (kms * 1e-3) / 1e9 / hbm_peak if kms else None
    # verifier chain of the Gruen proof: p_j(0) + p_j(1) = claim_j, claim_{j+1} = p_j(r_j), first claim = the input claim
    g_evals, g_chal, g_final_claim = res[0]
    from oracle import poly as OP

    cl = _fr_ints(claim_g)[0]
    ok = True
    for j in range(lg):
        e = _fr_ints(g_evals[j])
        ok = ok and (e[0] + e[1]) % B.FR_MOD == cl
        cl = OP.evaluate_cubic_at_point(e, _fr_ints(g_chal[j])[0])
    out[f"sumcheck_gruen_2^{lg}_ok"] = bool(ok and cl == _fr_ints(g_final_claim)[0])
    # Stage4GruenProver (registers read/write checking), K = 128 registers x T = 2^(sc_log_n - 8) cycles:
    # val, rd_wa, rs1_ra, rs2_ra = the first halves of the four resident tables (2^(sc_log_n - 1) entries each), inc = T entries
    # from the second half of the first.  Whole prover per run: ra, replicated inc, 7 + log T rounds with challenges from the
    # host, re-layout between the phases, final claims.
    if lg >= 12:
        log_T = lg - 8
        ptrs4 = d_tab_ptrs[:4] + [d_tab_ptrs[0] + (32 << (lg - 1))]
        gamma, r_cyc = rows_global(STREAM_EQ, 1 << 17, 1), rows_global(STREAM_EQ, (1 << 17) + 64, log_T)
        chal4 = rows_global(STREAM_EQ, (1 << 17) + 128, 7 + log_T)
        first = H.stage4_run(ctx, log_T, *ptrs4, gamma, r_cyc, -1, challenges=chal4)  # computes the input claim (untimed)
        claim4 = first[3][0].copy()

        def run_s4():
            res[0] = H.stage4_run(ctx, log_T, *ptrs4, gamma, r_cyc, -1, input_claim=claim4, challenges=chal4)

        run_s4()
        ctx.prof_enable(True)
        wall = time_ms(run_s4, 3, warm=0)
        ks = [ms for tag, ms in ctx.prof_read() if tag in (16, 17, 18)]
        ctx.prof_enable(False)
        per = len(ks) // 3
        p1 = log_T // 2
        k_p1 = sum(sum(ks[i * per: i * per + p1 + 1]) for i in range(3)) / 3
        n4 = 128 << log_T
        bytes_p1 = 32 * (4 * n4 + sum(6 * (n4 >> r) + 6 * (n4 >> (r + 1)) for r in range(p1)))
        out[f"stage4_gruen_T2^{log_T}_ms"] = wall
        out[f"stage4_gruen_T2^{log_T}_kernel_ms"] = sum(ks) / 3
        out[f"stage4_gruen_T2^{log_T}_phase1_frac_hbm"] = bytes_p1 / (k_p1 * 1e-3) / 1e9 / hbm_peak if k_p1 else None
        s_evals, s_chal, s_fin, s_claims = res[0]
        cl = _fr_ints(claim4)[0]
        ok = True
        for j in range(7 + log_T):
            e = _fr_ints(s_evals[j])
            ok = ok and (e[0] + e[1]) % B.FR_MOD == cl
            cl = OP.evaluate_cubic_at_point(e, _fr_ints(s_chal[j])[0])
        f = _fr_ints(s_fin)
        g1 = _fr_ints(gamma)[0]
        body = ((g1 * f[1] + g1 * g1 * f[2]) * f[0] + f[3] * (f[0] + f[4])) % B.FR_MOD
        out[f"stage4_gruen_T2^{log_T}_ok"] = bool(ok and cl == f[5] * body % B.FR_MOD and (s_evals == first[0]).all())
        # the Stage 3 provers (Shift / InstructionInput / RegistersClaimReduction, stage3_prover.zig) run together
        # like generateStage3Proof's loop over T = 2^(sc_log_n - 2) cycles; the 16 trace columns are the quarters of the four
        # resident tables.  Whole prover per run: eq / eq+1 tables, Q-buffer folds, batched columns, input claims, log T rounds
        # with challenges from the host, final claims.
        log_T3 = lg - 2
        cols3 = [d_tab_ptrs[t] + q * (32 << log_T3) for t in range(4) for q in range(4)]
        r_o, r_p = rows_global(STREAM_EQ, (1 << 17) + 256, log_T3), rows_global(STREAM_EQ, (1 << 17) + 320, log_T3)
        gam3 = rows_global(STREAM_EQ, (1 << 17) + 384, 7)
        g_int = _fr_ints(gam3[0:1])[0]
        gam3[0:5] = B.ints_to_mont_array([pow(g_int, i, B.FR_MOD) for i in range(5)], B.FR_MOD)
        bc3, chal3 = rows_global(STREAM_EQ, (1 << 17) + 448, 3), rows_global(STREAM_EQ, (1 << 17) + 512, log_T3)

        def run_s3():
            res[0] = H.stage3_run(ctx, log_T3, cols3, r_o, r_p, gam3, bc3, chal3)

        run_s3()
        ctx.prof_enable(True)
        wall = time_ms(run_s3, 3, warm=0)
        ks = [ms for tag, ms in ctx.prof_read() if tag in (16, 17, 18)]
        ctx.prof_enable(False)
        out[f"stage3_T2^{log_T3}_ms"] = wall
        out[f"stage3_T2^{log_T3}_round_kernel_ms"] = sum(ks) / 3
        # InstructionInputProver streams 2 x 5 tables, the witness binders 5 + 3: 128 B per entry and table over a whole proof
        out[f"stage3_T2^{log_T3}_rounds_frac_hbm"] = 128 * 18 * (1 << log_T3) / (sum(ks) / 3 * 1e-3) / 1e9 / hbm_peak if ks else None
        e3, c3, f3 = res[0]
        bcs, cl3 = _fr_ints(bc3), _fr_ints(c3)
        comb = sum(a * b for a, b in zip(cl3[:3], bcs)) % B.FR_MOD
        ok = True
        for j in range(log_T3):
            e = _fr_ints(e3[j])[10:14]
            ok = ok and (e[0] + e[1]) % B.FR_MOD == comb
            comb = OP.evaluate_cubic_at_point(e, _fr_ints(chal3[j:j + 1])[0])
        ok = ok and comb == sum(a * b for a, b in zip(cl3[3:], bcs)) % B.FR_MOD
        # final-claim relations (stage3_prover.zig:620-700)
        from oracle import stage3 as OS3

        rev = _fr_ints(chal3)[::-1]
        ro_i, rp_i, gi, fi = _fr_ints(r_o), _fr_ints(r_p), _fr_ints(gam3), _fr_ints(f3)
        Pm = B.FR_MOD
        sval = (fi[0] + gi[1] * fi[1] + gi[2] * fi[2] + gi[3] * fi[3]) % Pm
        ok = ok and cl3[3] == (OS3.eq_plus_one_mle(ro_i, rev) * sval + gi[4] * ((1 - fi[4]) % Pm) % Pm * OS3.eq_plus_one_mle(rp_i, rev)) % Pm
        left, right = (fi[5] * fi[6] + fi[7] * fi[8]) % Pm, (fi[9] * fi[10] + fi[11] * fi[12]) % Pm
        ok = ok and cl3[4] == (OP.eq_mle(ro_i, rev) + gi[5] * gi[5] % Pm * OP.eq_mle(rp_i, rev)) % Pm * ((right + gi[5] * left) % Pm) % Pm
        ok = ok and cl3[5] == OP.eq_mle(ro_i, rev) * ((fi[13] + gi[6] * fi[14] + gi[6] * gi[6] % Pm * fi[15]) % Pm) % Pm
        out[f"stage3_T2^{log_T3}_ok"] = bool(ok)
        # LassoProver (lasso/prover.zig) with the reference's own deriveChallenge: 2^(sc_log_n - 2) lookups with
        # 64-bit indices -- 64 one-pass address rounds over eq_evals and the index list, then the half-split cycle rounds
        log_TL, log_KL = lg - 2, 64
        idx_host = np.zeros((1 << log_TL, 2), dtype=np.uint64)
        idx_host[:, 0] = np.random.default_rng([0x5A17, 0x1A550]).integers(0, 1 << 64, size=1 << log_TL, dtype=np.uint64)
        d_idx = Z.DeviceBuffer.from_host(ctx, idx_host)
        r_red = rows_global(STREAM_EQ, (1 << 17) + 640, log_TL)

        def run_lasso():
            res[0] = H.lasso_run(ctx, d_idx.ptr, 1 << log_TL, log_TL, log_KL, r_red)

        run_lasso()
        wall = time_ms(run_lasso, 3, warm=0)
        d_idx.free()
        out[f"lasso_T2^{log_TL}_K2^{log_KL}_ms"] = wall
        lc, lcl, lch, lfin = res[0]
        cl_i, ch_i = _fr_ints(lcl), _fr_ints(lch)
        ok = cl_i[0] == 1  # every cycle has a lookup: eq(r, .) sums to one
        fe = 1
        for j in range(log_TL + log_KL):
            c = _fr_ints(lc[j])
            ok = ok and (2 * c[0] + c[1] + c[2]) % B.FR_MOD == cl_i[j] and (c[0] + c[1] * ch_i[j] + c[2] * ch_i[j] * ch_i[j]) % B.FR_MOD == cl_i[j + 1]
            if j < log_KL:
                fe = fe * (1 - ch_i[j]) % B.FR_MOD
        out[f"lasso_T2^{log_TL}_K2^{log_KL}_ok"] = bool(ok and _fr_ints(lfin.reshape(1, 4))[0] == fe)
    # K4: Dory row commitments (dory.zig:989-1042) of a 2^log_n-entry polynomial: 2^nu row MSMs over the first 2^sigma bases,
    # device-resident evaluations (the MSM scalars); rows 0 and last checked against the closed form (sum_j e[r, j] k_j) G
    d_ev = Z.DeviceBuffer.from_host(ctx, h_scalars)
    try:
        def run_dory():
            res[0] = Z.Dory.row_commitments(bases, d_ev.ptr, device_len=n)

        run_dory()
        out["dory_rows_ms"] = time_ms(run_dory, 3, warm=0)
        xy, inf, rows_d, cols_d = res[0]
        out["dory_rows_shape"] = f"{rows_d}x{cols_d}"
        ok = True
        for r_ in (0, rows_d - 1):
            ok = ok and check_msm_closed_form((xy[r_], inf[r_]), [R.fr_dot(h_logs[:cols_d], h_scalars[r_ * cols_d:(r_ + 1) * cols_d])])
        out["dory_rows_ok"] = bool(ok)
        # the same rows over the G1 vector as a long-lived SRS handle of its own (DorySRS.g1_vec, 2^sigma points) with its
        # window-level table (c = 12, built untimed like the headline MSM's table)
        g1 = Z.Bases.fixed_base_mul(ctx, h_logs[:cols_d])
        try:
            g1.precompute(12)

            def run_dory_t():
                res[0] = Z.Dory.row_commitments(g1, d_ev.ptr, device_len=n)

            run_dory_t()
            out["dory_rows_table_ms"] = time_ms(run_dory_t, 3, warm=0)
            xy_t, inf_t, _, _ = res[0]
            out["dory_rows_table_ok"] = bool((xy_t == xy).all() and (inf_t == inf).all())
        finally:
            g1.free()
        # f1: HyperKZG.open (commitment/mod.zig:261-324) of the same polynomial at a random point: log n quotient commitments
        # (MSMs of 2^(log n - 1) ... 1 points over the resident bases) and the folds between them, nothing leaves the device
        # but the commitments; the final evaluation is checked against a chain of zolt_poly_bind_first over the same point
        log_n_open = n.bit_length() - 1
        pt = rows_global(STREAM_EQ, (1 << 17) + 768, log_n_open)

        def run_open():
            res[0] = Z.HyperKZG.open_dev(bases, d_ev.ptr, n, pt)

        run_open()
        out[f"hyperkzg_open_2^{log_n_open}_ms"] = time_ms(run_open, 3, warm=0)
        q_xy, q_inf, fin = res[0]
        d_tmp = Z.DeviceBuffer(ctx, 32 * (n // 2 + 1))
        try:
            cur, cur_n = d_ev.ptr, n
            for j in range(log_n_open):  # open's own fold, (1 - r_i) lo + r_i hi over the two halves (mod.zig:298-305) = bindFirst
                rj = np.ascontiguousarray(pt[j])
                Z._lib.check(ctx.lib.zolt_poly_bind_first(ctx.handle, C.c_void_p(d_tmp.ptr), C.c_void_p(cur), cur_n,
                                                          rj.ctypes.data_as(C.c_void_p)), "zolt_poly_bind_first")
                cur, cur_n = d_tmp.ptr, cur_n // 2
            ev0 = d_tmp.download((1, 4))
        finally:
            d_tmp.free()
        out[f"hyperkzg_open_2^{log_n_open}_quotients"] = int(len(q_xy))
        out[f"hyperkzg_open_2^{log_n_open}_final_matches_bind_chain"] = bool((np.asarray(fin).reshape(-1)[:4] == ev0.reshape(-1)).all())
    finally:
        d_ev.free()
    return {k: (round(v, 5) if isinstance(v, float) else v) for k, v in out.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log-n", type=int, default=24, help="MSM size (log2 points)")
    ap.add_argument("--sc-log-n", type=int, default=24, help="sumcheck table size (log2 entries)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the single-GPU extra legs (configs 1, 2, Gruen, trace-like)")
    ap.add_argument("--nccl-exchange", action="store_true", help="multi-GPU: host-driven NCCL all-gathers instead of the NVLink mailboxes")
    ap.add_argument("--no-table", action="store_true", help="MSM without the precomputed window-level table")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
