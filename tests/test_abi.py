"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports exactly the entry
points include/kb2e_b200.h declares; no compute is attempted without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "kb2e_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kb2e_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_documented_entry_points():
    syms = declared_symbols()
    for s in ("kb2e_create", "kb2e_destroy", "kb2e_train_epochs", "kb2e_rank", "kb2e_score", "kb2e_upload", "kb2e_download"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    import kb2e_b200
    from kb2e_b200 import api
    lib = kb2e_b200.load_library()
    for s in declared_symbols():
        assert hasattr(lib, s), f"libkb2e_b200.so does not export {s}"
    assert sorted(api.SYMBOLS) == declared_symbols()


def test_create_fails_loudly_without_a_gpu():
    """No CPU fallback: on a box without a CUDA device kb2e_create must return an error."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import kb2e_b200
    with pytest.raises(kb2e_b200.Kb2eError) as e:
        kb2e_b200.Context("transe", 8, 10, 2)
    assert "no CUDA device" in str(e.value) or "CPU fallback" in str(e.value)


def test_bad_config_is_rejected_before_touching_the_device():
    import kb2e_b200
    lib = kb2e_b200.load_library()
    from kb2e_b200.api import Config
    cfg = Config(7, 8, 0, 0, 1, 0, 10, 2, 0.01, 1.0, 0, 0, 0)
    ptr = ctypes.c_void_p()
    assert lib.kb2e_create(ctypes.byref(cfg), ctypes.byref(ptr)) == 1
    assert b"unknown model" in lib.kb2e_last_error(None)
    assert lib.kb2e_create(None, ctypes.byref(ptr)) == 1


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under kb2e_b200/ may reference it."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "kb2e_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".cc", ".h", "Makefile")):
                text = open(os.path.join(base, f), errors="ignore").read()
                # comments may name the oracle twin of a kernel; code may not include / import / load it
                if re.search(r"(#\s*include|import|from|CDLL|dlopen)[^\n]*(oracle|_ref)|libkb2e_(ref|oracle)|/root/reference", text):
                    bad.append(os.path.join(base, f))
    assert not bad, bad
