"""CPU-only: libtdsfs.so loads and exports every symbol include/tdsfs.h declares; no compute without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "tdsfs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tdsfs_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    import tdsfs_capi
    lib = ctypes.CDLL(tdsfs_capi.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/tdsfs.h but not exported"
    assert sorted(tdsfs_capi.EXPORTS) == syms


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    import tdsfs_capi
    with pytest.raises(tdsfs_capi.TdsfsError):
        tdsfs_capi.Handle(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "2dsfs-scan_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert "sfs_oracle" not in txt and "oracle/" not in txt, f"{fn} references the oracle"
