"""The C-ABI library loads and exports every symbol include/owc_b200.h declares (no GPU compute)."""

import ctypes
import os
import re

import pytest
import torch

import owc_b200  # noqa: F401
from owc_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(headers=("owc_b200.h", "owc_b200_debug.h")):
    names = set()
    for h in headers:
        text = open(os.path.join(ROOT, "include", h), encoding="utf-8").read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(owc_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_product_header_carries_no_debug_hooks():
    product = _declared_symbols(("owc_b200.h",))
    assert not [n for n in product if "debug" in n or "bench" in n]
    assert "owc_match" in product and "owc_score" in product


def test_header_symbols_are_exported_and_bound():
    names = _declared_symbols()
    assert len(names) >= 12
    handle = ctypes.CDLL(_cabi.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in the header but not exported by the library"
    assert sorted(_cabi.EXPORTS) == names, "ctypes table and header disagree"


def test_abi_version_and_error_string():
    lib = _cabi.lib()
    assert lib.owc_abi_version() == _cabi.ABI_VERSION
    assert isinstance(lib.owc_last_error(), bytes)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly():
    lib = _cabi.lib()
    ctx = ctypes.c_void_p()
    rc = lib.owc_ctx_create(0, None, 0, ctypes.byref(ctx))
    assert rc != 0 and not ctx
    assert b"no CPU fallback" in lib.owc_last_error() or b"CUDA" in lib.owc_last_error()
    from owc_b200 import arch as A
    from owc_b200.engine import Engine

    with pytest.raises(_cabi.OwcError):
        Engine(A.MINILM_L6, {}, device=0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "lmms-owc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
