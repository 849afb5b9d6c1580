"""The C-ABI library loads and exports every symbol include/pa_engine.h declares
(no compute calls: runs without a GPU)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pa_engine.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pa_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_functions():
    fns = declared_functions()
    assert "pa_ctx_create" in fns and "pa_fixed_base_mul" in fns and len(fns) >= 20


def test_library_exports_every_declared_symbol(pa):
    lib = pa.load_library()
    for name in declared_functions():
        assert hasattr(lib, name), f"libpa_engine.so does not export {name}"


def test_binding_covers_header(pa):
    assert sorted(pa.SIGNATURES) == declared_functions()


def test_library_is_sm100a_only(pa):
    out = subprocess.run(["cuobjdump", "--list-elf", pa.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_gpu_means_loud_failure(pa):
    """Without a CUDA device the engine must refuse to construct (no CPU path)."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(pa.EngineError):
        pa.Engine(0)


def test_product_does_not_reference_oracle():
    """Nothing under the product package or include/ may mention the oracle."""
    bad = []
    for base in ("privacy-auction_b200", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".so", ".o", ".log", ".pyc")):
                    continue
                txt = open(os.path.join(d, f), errors="ignore").read()
                if "pa_oracle" in txt or "oracle/" in txt or "libcrypto.so" in txt:
                    bad.append(os.path.join(d, f))
    assert not bad, bad


def test_struct_layouts_match_the_binding(pa):
    """the ctypes mirrors of the job structs must have the library's sizes (no GPU needed)"""
    import ctypes
    import sys
    eng = sys.modules[pa.Engine.__module__]
    lib = pa.load_library()
    assert lib.pa_abi_sizeof(0) == ctypes.sizeof(eng.SealJob)
    assert lib.pa_abi_sizeof(1) == ctypes.sizeof(eng.Ccs22Job)
    assert lib.pa_abi_sizeof(2) == ctypes.sizeof(eng.KernelStat)
    assert lib.pa_abi_sizeof(9) == 0 and lib.pa_abi_version() == 3
