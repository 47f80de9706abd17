"""ctypes access to the scalar C oracle (``oracle/wigner_oracle.c``).

ORACLE -- test infrastructure, not product code (see ``oracle/__init__.py``).
Rows are split over Python threads (ctypes drops the GIL), so the C side stays
free of any threading runtime.
"""

from __future__ import annotations

import ctypes
import math
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(
        os.path.join(_HERE, "wigner_oracle.c")
    ):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        dp = ctypes.POINTER(ctypes.c_double)
        for name in ("oracle_wigner_clenshaw_f64", "oracle_wigner_clenshaw_ld"):
            fn = getattr(L, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [dp, ctypes.c_int, ctypes.c_int64, dp, ctypes.c_int, dp, ctypes.c_int,
                           ctypes.c_double, ctypes.c_int, ctypes.c_int, dp]
        L.oracle_partial_trace_sv.restype = ctypes.c_int
        L.oracle_partial_trace_sv.argtypes = [dp, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int, dp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def wigner_clenshaw(rho, xvec, yvec=None, g: float = math.sqrt(2), long_double: bool = False,
                    threads: int | None = None) -> np.ndarray:
    rho = np.ascontiguousarray(np.asarray(getattr(rho, "data", rho), dtype=np.complex128))
    N = rho.shape[0]
    xvec = np.ascontiguousarray(xvec, dtype=np.float64)
    yvec = xvec if yvec is None else np.ascontiguousarray(yvec, dtype=np.float64)
    nx, ny = len(xvec), len(yvec)
    W = np.empty((ny, nx), dtype=np.float64)
    fn = lib().oracle_wigner_clenshaw_ld if long_double else lib().oracle_wigner_clenshaw_f64
    rv = rho.view(np.float64)
    T = max(1, min(threads or os.cpu_count() or 1, ny))
    cuts = [ny * t // T for t in range(T + 1)]

    def run(t):
        return fn(_dp(rv), N, N, _dp(xvec), nx, _dp(yvec), ny, float(g), cuts[t], cuts[t + 1], _dp(W))

    if T == 1:
        rcs = [run(0)]
    else:
        with ThreadPoolExecutor(T) as ex:
            rcs = list(ex.map(run, range(T)))
    if any(rcs):
        raise RuntimeError(f"C oracle failed: {rcs}")
    return W


def partial_trace_sv(psi, qargs) -> np.ndarray:
    psi = np.ascontiguousarray(np.asarray(psi, dtype=np.complex128).reshape(-1))
    n = int(psi.shape[0]).bit_length() - 1
    q = np.ascontiguousarray(list(qargs), dtype=np.int32)
    D = 1 << (n - len(q))
    out = np.empty((D, D), dtype=np.complex128)
    rc = lib().oracle_partial_trace_sv(_dp(psi.view(np.float64)), n,
                                       q.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), len(q),
                                       _dp(out.view(np.float64)))
    if rc:
        raise RuntimeError(f"C oracle partial trace failed: {rc}")
    return out
