"""The C-ABI library loads and exports every symbol include/bq_b200.h declares (no compute, no GPU)."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bq_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bq_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_loads():
    from bosonic_qiskit_b200 import _build, _lib

    path = _build.build_library()
    assert os.path.exists(path)
    lib = _lib.load()
    assert lib.bq_version() == 100


def test_every_declared_symbol_is_exported_and_bound():
    from bosonic_qiskit_b200 import _lib

    lib = _lib.load()
    names = declared_functions()
    assert len(names) >= 30
    raw = ctypes.CDLL(_lib.library_path())
    for name in names:
        assert hasattr(raw, name), f"{name} declared in bq_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes prototype in _lib.SIGNATURES"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"


def test_header_cites_reference_seams():
    src = open(HEADER).read()
    for cite in ("wigner.py:190", "util.py:580,596,619,693", "wigner.py:92-96", "animate.py:191-206", "util.py:688-694"):
        assert cite in src


def test_header_compiles_as_c():
    import subprocess
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write('#include "bq_b200.h"\nint (*fp)(void) = bq_version;\nint main(void){ return fp == 0; }\n')
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", c, "-o",
                        os.path.join(d, "t.o")], check=True)


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product path must fail loudly, not compute on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from bosonic_qiskit_b200 import BQError, _lib, get_engine

    h = ctypes.c_void_p()
    rc = _lib.load().bq_create(ctypes.byref(h), 0)
    assert rc == _lib.BQ_ERR_NO_DEVICE and not h.value
    assert b"no CPU fallback" in _lib.load().bq_last_error()
    with pytest.raises(BQError):
        get_engine(0)
    import numpy as np

    from bosonic_qiskit_b200 import wigner

    with pytest.raises(BQError):
        wigner.wigner(np.eye(4) / 4)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "bosonic_qiskit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "oracle/" not in text and "liboracle" not in text, f
