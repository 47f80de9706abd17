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
    assert lib.bq_version() == 200


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


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` needs no GPU: it must run here, print exactly one JSON line with the contract's
    keys, and (under torchrun) only on rank 0."""
    import json
    import subprocess
    import sys

    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c1", "--steps", "1", "--warmup", "1",
           "--cpu-seconds", "1"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] in ("port", "reference")
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run(cmd + ["--gpus", "2"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    res = subprocess.run(cmd[:4] + ["--config", "c5", "--steps", "1", "--warmup", "1", "--cpu-seconds", "1", "--gpus", "2"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    assert json.loads(res.stdout.strip())["scaling"] == "strong"
