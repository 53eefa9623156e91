"""CPU: the C-ABI library loads without a GPU and exports exactly what include/emia.h declares; the product path refuses to
run without CUDA (no CPU fallback); nothing in the product package imports the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "emia.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(emia_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    import __graft_entry__ as ge
    ge.build()
    from deepemia_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/emia.h but not exported by libemia.so"
    assert sorted(_lib.EXPORTS) == names, "ctypes signature table and header disagree"
    assert _lib.load().emia_version() >= 100


def test_bad_arguments_return_error_codes():
    from deepemia_b200 import _lib
    lib = _lib.load()
    assert lib.emia_paste_plan(None, 4, 1.0, 1.0, 16, 16, None, None, None) == -1
    assert b"emia_paste_plan" in lib.emia_last_error()
    assert lib.emia_paste_threshold_bitpack(None, None, None, None, -1, 1.0, 1.0, 8, 8, None, 1, 4, None, None, None, 0, None, None) == -1
    assert lib.emia_dedup_smart(*([None] * 10), -1, 0, 0, None, None, 0.5, 0.0, None, None, None, 0, None) == -1
    assert lib.emia_exclusive_scan_i64(None, 0, None, 0, None) == -1


def test_product_path_has_no_cpu_fallback_and_no_oracle_import():
    import torch
    from deepemia_b200 import _lib, engine
    if not torch.cuda.is_available():
        with pytest.raises(_lib.EmiaError):
            engine._need_cuda(None)
    pkg = os.path.join(ROOT, "deepemia_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"
