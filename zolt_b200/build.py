"""Build the CUDA C-ABI library in-tree for sm_100a.

    python -m zolt_b200.build            # build if sources are newer than the library
    python -m zolt_b200.build --force

Outputs (git-ignored, shipped to the GPU box by gpurun):
    zolt_b200/lib/libzolt_gpu.so   the product: CUDA kernels + C ABI (include/zolt_gpu.h)
    zolt_b200/lib/libzolt_gpu.a    static archive for `exe.addObjectFile` in build.zig
    zolt_b200/lib/libzolt_host.so  C++ host mirror of the reference interfaces (zolt_b200/host)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
HOST = ROOT / "host"
LIB = ROOT / "lib"
OBJ = LIB / "obj"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU path)")
    return cand


def _newer(src_files, target: Path) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in src_files)


def _run(cmd, log: Path | None = None):
    res = subprocess.run(cmd, capture_output=True, text=True)
    if log is not None:
        log.write_text(res.stdout + res.stderr)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("command failed: " + " ".join(map(str, cmd)))
    return res


def build_gpu(force: bool = False, verbose: bool = False) -> Path:
    LIB.mkdir(exist_ok=True)
    OBJ.mkdir(exist_ok=True)
    nvcc = _nvcc()
    headers = list(CSRC.glob("*.cuh")) + [ROOT.parent / "include" / "zolt_gpu.h"]
    units = sorted(CSRC.glob("*.cu"))
    so = LIB / "libzolt_gpu.so"

    def compile_one(cu: Path) -> Path:
        obj = OBJ / (cu.stem + ".o")
        if force or _newer([cu, *headers], obj):
            if verbose:
                print(f"[zolt_b200.build] nvcc {cu.name}", flush=True)
            _run([nvcc, *NVCC_FLAGS, "-c", str(cu), "-o", str(obj)], log=OBJ / (cu.stem + ".ptxas.log"))
        return obj

    with ThreadPoolExecutor(max_workers=len(units)) as ex:
        objs = list(ex.map(compile_one, units))
    if force or _newer(objs, so):
        _run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *map(str, objs), "-o", str(so)])
        ar = LIB / "libzolt_gpu.a"
        if ar.exists():
            ar.unlink()
        _run(["ar", "rcs", str(ar), *map(str, objs)])
    return so


def build_host(force: bool = False, verbose: bool = False) -> Path | None:
    srcs = sorted(HOST.glob("*.cpp"))
    if not srcs:
        return None
    so = LIB / "libzolt_host.so"
    deps = srcs + list(HOST.glob("*.hpp")) + [ROOT.parent / "include" / "zolt_gpu.h"]
    if force or _newer(deps, so):
        if verbose:
            print("[zolt_b200.build] g++ host mirror", flush=True)
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", str(ROOT.parent / "include"),
              *map(str, srcs), "-o", str(so), "-L", str(LIB), "-lzolt_gpu", "-Wl,-rpath,$ORIGIN"])
    return so


def build_all(force: bool = False, verbose: bool = False):
    gpu = build_gpu(force, verbose)
    host = build_host(force, verbose)
    return gpu, host


if __name__ == "__main__":
    out = build_all(force="--force" in sys.argv, verbose=True)
    print("built:", *[str(o) for o in out if o])
