"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/pfn_b200.h declares."""
import ctypes
import os
import re

from transformerscandobayesianinference_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pfn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pfn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    lib = L.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/pfn_b200.h but not exported"
    assert set(L.EXPORTED_SYMBOLS) == set(declared)


def test_version_and_error_string():
    lib = L.load()
    assert lib.pfn_version() == 1
    d = L.GemmDesc()
    rc = lib.pfn_gemm_simt(ctypes.byref(d), None)       # argument validation happens before any CUDA call
    assert rc != 0 and b"empty problem" in lib.pfn_last_error()
    a = L.AttnDesc()
    assert lib.pfn_attention_fwd_tc(ctypes.byref(a), None) != 0
