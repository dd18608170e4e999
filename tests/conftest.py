"""pytest configuration.

Tiers
  -m "not gpu" : oracle vs the reference's golden vectors, host logic, C-ABI symbol checks, host-shim
                 validation of the device arithmetic, world-size-2 gloo tests.  Runs without a GPU.
  -m gpu       : parity tests proper -- the CUDA path, called through the C ABI, against the oracle.
"""
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def built_lib():
    """Path of libzolt_gpu.so, building it when nvcc is available and the library is stale."""
    from zolt_b200 import build

    import shutil

    so = build.LIB / "libzolt_gpu.so"
    have_nvcc = bool(shutil.which("nvcc")) or os.path.exists("/usr/local/cuda/bin/nvcc")
    if not have_nvcc:
        # the only tolerated case: a box without the toolkit runs the library that travelled with the snapshot
        if not so.exists():
            pytest.fail("libzolt_gpu.so is missing and nvcc is not installed: nothing to test (there is no CPU path)")
        return so
    try:
        build.build_all()
    except Exception as e:  # noqa: BLE001 - a broken source tree must never be tested against a stale binary
        pytest.fail(f"libzolt_gpu.so could not be rebuilt from the current sources: {e}")
    return so


@pytest.fixture(scope="session")
def ctx(built_lib):
    import zolt_b200

    c = zolt_b200.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def host_shim():
    """ctypes handle on tests/host_shim (device headers compiled with the PTX-emulation back-end)."""
    import ctypes
    import subprocess

    src = ROOT / "tests" / "host_shim" / "host_shim.cpp"
    out = ROOT / "tests" / "host_shim" / "libhost_shim.so"
    deps = [src] + list((ROOT / "zolt_b200" / "csrc").glob("*.cuh"))
    if not out.exists() or any(d.stat().st_mtime > out.stat().st_mtime for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(out), str(src)], check=True)
    return ctypes.CDLL(str(out))
