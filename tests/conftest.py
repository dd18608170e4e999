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

    try:
        build.build_all()
    except Exception as e:  # noqa: BLE001 - surface the reason in the skip/fail message
        if not build.LIB.joinpath("libzolt_gpu.so").exists():
            pytest.fail(f"libzolt_gpu.so is missing and could not be built: {e}")
    return build.LIB / "libzolt_gpu.so"


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
