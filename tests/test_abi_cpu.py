"""CPU tests (no GPU): the C-ABI shared library loads, exports every symbol include/flipb200.h
declares, and fails loudly (no CPU fallback) when there is no device. No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "flipb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(flipb200_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    from zeno_b200 import abi
    syms = header_symbols()
    assert len(syms) >= 30
    assert sorted(abi.EXPORTS) == syms


def test_library_exports_every_declared_symbol():
    from zeno_b200 import abi
    lib = abi.load_library()
    for s in header_symbols():
        assert hasattr(lib, s), f"libflipb200.so does not export {s}"
    assert lib.flipb200_abi_version() == 1
    info = lib.flipb200_build_info().decode()
    assert "sm_100a" in info


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path needs a CPU-only box")
    from zeno_b200 import abi
    with pytest.raises(abi.FlipB200Error) as e:
        abi.World(0.1)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "zeno_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".sh")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in text and "liboracle" not in text and "flip_oracle.h" not in text, f


def test_cuda_sources_are_sm100a_native():
    sh = open(os.path.join(ROOT, "zeno_b200", "csrc", "build.sh")).read()
    assert "arch=compute_100a,code=sm_100a" in sh and "-lineinfo" in sh
