"""CPU tier: the C-ABI library loads and exports every function include/zolt_gpu.h declares; with no
GPU the product fails loudly instead of falling back to a CPU path."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_functions():
    text = (ROOT / "include" / "zolt_gpu.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zolt_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(built_lib):
    lib = ctypes.CDLL(str(built_lib))
    names = _declared_functions()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_table_matches_header(built_lib):
    from zolt_b200 import _lib

    assert sorted(list(_lib.SIGNATURES) + ["zolt_gpu_version"]) == _declared_functions()
    _lib.load_library()  # attaches every prototype; AttributeError if a symbol is missing


def test_no_gpu_means_loud_failure(built_lib):
    import torch

    import zolt_b200

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(zolt_b200.ZoltGpuError) as ei:
        zolt_b200.Context(0)
    assert ei.value.status == 4  # ZOLT_ERR_NOGPU


def test_product_never_imports_oracle():
    # the oracle is test infrastructure: nothing under zolt_b200/ may reference it
    for path in list((ROOT / "zolt_b200").rglob("*.py")) + list((ROOT / "zolt_b200").rglob("*.c*")) + list((ROOT / "zolt_b200").rglob("*.hpp")):
        if "lib/" in str(path):
            continue
        text = path.read_text(errors="replace")
        assert "oracle" not in text.replace("no CPU", ""), path
