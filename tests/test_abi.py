"""The C-ABI library builds, loads without a GPU, and exports every symbol the header declares."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "aukit_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aukit_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(ak):
    lib = ctypes.CDLL(ak.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 45
    for n in names:
        assert hasattr(lib, n), "libaukit_cuda.so does not export %s" % n
    assert sorted(ak.SIGNATURES) == names          # the ctypes table binds exactly the header


def test_library_does_not_link_oracle_or_torch(ak):
    out = subprocess.run(["ldd", ak.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "torch" not in out
    syms = subprocess.run(["nm", "-D", ak.LIB_PATH], capture_output=True, text=True).stdout
    assert "auko_" not in syms


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "aukit_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h", ".lua")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.lower() or f == "__init__.py" and False, "%s mentions the oracle" % f


def test_fails_loudly_without_gpu(ak):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ak.AukitError, match="no CPU fallback"):
        ak.pcm(b"\0\0", 16)


def test_abi_version(ak):
    assert ctypes.CDLL(ak.LIB_PATH).aukit_cuda_abi_version() == 1
