"""Thin object wrapper over the C ABI: one ``Engine`` per GPU.

NumPy in / NumPy out goes through the ``*_host`` entry points (host buffers,
staging inside the library); ``torch`` CUDA tensors go through the ``*_dev``
entry points and stay on the device.  torch is imported lazily and only used
for device memory and streams.
"""

from __future__ import annotations

import ctypes
import math
import threading
import weakref
from typing import Sequence

import numpy as np

from . import _lib
from ._lib import BQError, check

METHODS = {"clenshaw": 0, "iterative": 1, "laguerre": 2, "fft": 3}


def method_code(method: str) -> int:
    try:
        return METHODS[method]
    except (KeyError, TypeError):
        # same exception type / text qutip.wigner raises for an unknown method
        raise TypeError("method must be either 'iterative', 'laguerre', 'fft' or 'clenshaw'.") from None


def _c128(a, copy=False) -> np.ndarray:
    out = np.ascontiguousarray(a, dtype=np.complex128)
    return out.copy() if copy and out is a else out


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a: np.ndarray):
    return ctypes.c_void_p(a.ctypes.data)


def _int_array(v: Sequence[int]):
    arr = (ctypes.c_int * max(len(v), 1))(*[int(x) for x in v])
    return arr


def _n_qubits(dim: int) -> int:
    n = int(dim).bit_length() - 1
    if dim <= 0 or (1 << n) != dim:
        raise ValueError(f"dimension {dim} is not a power of two")
    return n


class Engine:
    """Owns a ``bq_handle_t`` on one CUDA device."""

    PINNED_POOL_LIMIT = 2 << 30  # bytes of recycled pinned host buffers kept per engine

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        check(self._lib.bq_create(ctypes.byref(h), int(device)))
        self._h = h
        self.device = int(device)
        self._finalizer = weakref.finalize(self, self._lib.bq_destroy, h)
        self._pool: dict[int, list[int]] = {}
        self._pool_bytes = 0
        self._pool_lock = threading.Lock()
        self._grid_lock = threading.Lock()
        self._grid_sharing = True
        self._grid_verdicts: dict[int, tuple] = {}
        self._stage = None  # local block for symmetric-grid results bound for a peer buffer

    def close(self) -> None:
        self._finalizer()

    # ------------------------------------------------------------------ bookkeeping
    @property
    def launches(self) -> int:
        n = ctypes.c_uint64()
        check(self._lib.bq_launch_count(self._h, ctypes.byref(n)))
        return int(n.value)

    def synchronize(self) -> None:
        check(self._lib.bq_synchronize(self._h))

    def last_device_ms(self) -> float:
        ms = ctypes.c_double()
        check(self._lib.bq_last_device_ms(self._h, ctypes.byref(ms)))
        return float(ms.value)

    def fp64_peak_tflops(self, seconds: float = 0.5) -> float:
        v = ctypes.c_double()
        check(self._lib.bq_fp64_peak(self._h, float(seconds), ctypes.byref(v)))
        return float(v.value)

    def set_sass_mode(self, patched: bool) -> bool:
        """Launch the re-scheduled SASS of the Wigner kernels (True) or nvcc's own schedule (False).
        Returns what is in effect (False when the build carries no patched image)."""
        active = ctypes.c_int()
        check(self._lib.bq_set_sass_mode(self._h, 1 if patched else 0, ctypes.byref(active)))
        return bool(active.value)

    def sass_self_check(self) -> tuple[int, int]:
        """(mask of kernel variants whose re-scheduled SASS failed the load-time bit-identity check, variants checked)."""
        bad, n = ctypes.c_uint64(), ctypes.c_int()
        check(self._lib.bq_sass_self_check(self._h, ctypes.byref(bad), ctypes.byref(n)))
        return int(bad.value), int(n.value)

    def set_grid_sharing(self, on: bool) -> None:
        """Mirror-symmetric grids (``yvec == xvec == -xvec[::-1]``, what the reference always builds): share the
        inner Clenshaw recurrence between the eight points of equal ``x**2 + p**2`` (default on).  Off = always the
        point-by-point kernels."""
        with self._grid_lock:
            self._grid_sharing = bool(on)
            check(self._lib.bq_set_grid_symmetry(self._h, 1 if on else 0))

    def set_sym_variant(self, variant: int | None) -> None:
        """Tuning / test aid: only this variant of the symmetric-grid kernels (None: the library's own choice)."""
        check(self._lib.bq_set_sym_variant(self._h, -1 if variant is None else int(variant)))

    # variant ids of csrc/wigner_kernels.cu (kKernelNames)
    KERNEL_VARIANTS = {
        0: "wigner_resident_kernel<4,256,2>", 1: "wigner_resident_kernel<2,128,4>", 2: "wigner_resident_kernel<1,128,4>",
        3: "wigner_stream_kernel<4,256,4,256>", 4: "wigner_stream_kernel<2,128,4,256>", 5: "wigner_resident_kernel<8,128,2>",
        6: "wigner_stream_kernel<8,128,4,256>", 7: "wigner_diag_kernel<8,256,3>", 8: "wigner_sym_resident_kernel<3,128,2>",
        9: "wigner_sym_resident_kernel<1,128,4>", 10: "wigner_sym_diag_kernel<3,256,3>", 11: "wigner_sym_resident_kernel<2,128,2>",
        12: "wigner_sym_resident_kernel<2,192,2>", 13: "wigner_sym_resident_kernel<1,256,2>", 14: "wigner_sym_diag_kernel<2,256,3>",
        15: "wigner_sym_diag_kernel<2,384,3>", 16: "wigner_sym_diag_kernel<1,256,3>", 17: "wigner_sym_resident_kernel<4,256,1>",
        18: "wigner_sym_diag_kernel<4,256,3>", 19: "wigner_sym_diag_parked_kernel<4,256,2>",
        20: "wigner_sym_diag_parked_kernel<3,384,2>", 21: "wigner_sym_diag_parked_kernel<4,256,2,trip8>",
        22: "wigner_sym_resident_mixed_kernel<4,3,256>",
    }

    def last_wigner_kernel(self) -> str:
        """Name of the kernel variant the last Wigner call launched."""
        v = ctypes.c_int()
        check(self._lib.bq_last_wigner_variant(self._h, ctypes.byref(v)))
        return self.KERNEL_VARIANTS.get(int(v.value), f"variant {int(v.value)}")

    def grid_is_mirror(self, xvec, yvec=None) -> bool:
        """The library's own test of a host grid (``bq_grid_is_mirror``)."""
        x = _f64(xvec).reshape(-1)
        y = x if yvec is None else _f64(yvec).reshape(-1)
        return bool(self._lib.bq_grid_is_mirror(_ptr(x), len(x), _ptr(y), len(y)))

    def _device_grid_mode(self, xvec, yvec) -> int:
        """Grid-symmetry mode for a ``_dev`` call: the library cannot read a device grid without a synchronisation,
        so the binding tests each grid TENSOR OBJECT once.  The verdict is tied to the object's lifetime (a weak
        reference evicts it) and to torch's version counter -- never to the storage address, which the caching
        allocator hands to the next tensor."""
        if not self._grid_sharing:
            return 0
        if yvec is not xvec and (yvec.data_ptr() != xvec.data_ptr() or yvec.numel() != xvec.numel()):
            if yvec.numel() != xvec.numel():
                return 1
            key_obj, other = xvec, yvec
        else:
            key_obj, other = xvec, None
        entry = self._grid_verdicts.get(id(key_obj))
        stamp = (key_obj._version, None if other is None else (id(other), other._version))
        if entry is not None and entry[0]() is key_obj and entry[1] == stamp and (other is None or entry[2]() is other):
            return 2 if entry[3] else 1
        xh = key_obj.detach().cpu().numpy()
        yh = xh if other is None else other.detach().cpu().numpy()
        verdict = self.grid_is_mirror(xh, yh)
        verdicts, kid = self._grid_verdicts, id(key_obj)
        if len(verdicts) > 256:
            verdicts.clear()

        def _evict(_ref, verdicts=verdicts, kid=kid):
            verdicts.pop(kid, None)

        verdicts[kid] = (weakref.ref(key_obj, _evict), stamp, None if other is None else weakref.ref(other), verdict)
        return 2 if verdict else 1

    def profile(self, on: bool) -> None:
        """Bracket every Wigner-kernel launch with CUDA events (bench.py's roofline leg)."""
        check(self._lib.bq_profile_enable(self._h, 1 if on else 0))

    def profile_read(self) -> tuple[float, int]:
        """(summed kernel ms, launches) since the last read."""
        ms, n = ctypes.c_double(), ctypes.c_int()
        check(self._lib.bq_profile_read(self._h, ctypes.byref(ms), ctypes.byref(n)))
        return float(ms.value), int(n.value)

    def pinned_empty(self, shape, dtype=np.float64) -> np.ndarray:
        """NumPy array in page-locked, device-mapped host memory.

        Outputs placed here are written by the GPU directly (no staging copy).  Buffers come from a
        small per-engine pool and return to it when the array is garbage collected, because pinning
        pages costs far more than the kernels this library runs.
        """
        dtype = np.dtype(dtype)
        count = int(np.prod(shape, dtype=np.int64))
        nbytes = count * dtype.itemsize
        cap = max(65536, -(-nbytes // 65536) * 65536)
        ptr = None
        with self._pool_lock:
            free = self._pool.get(cap)
            if free:
                ptr = free.pop()
                self._pool_bytes -= cap
        if ptr is None:
            p = ctypes.c_void_p()
            check(self._lib.bq_host_alloc(self._h, cap, ctypes.byref(p)))
            ptr = int(p.value)
        buf = (ctypes.c_char * cap).from_address(ptr)
        arr = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
        weakref.finalize(buf, Engine._recycle, weakref.ref(self), cap, ptr)
        return arr

    @staticmethod
    def _recycle(self_ref, cap: int, ptr: int) -> None:
        self = self_ref()
        if self is None or not self._finalizer.alive:
            return  # engine gone: bq_destroy does not own these, the process is exiting
        with self._pool_lock:
            if self._pool_bytes + cap <= self.PINNED_POOL_LIMIT:
                self._pool.setdefault(cap, []).append(ptr)
                self._pool_bytes += cap
                return
        self._lib.bq_host_free(self._h, ctypes.c_void_p(ptr))

    # ------------------------------------------------------------------ partial trace (host)
    def partial_trace_sv(self, psi, traced: Sequence[int]) -> np.ndarray:
        """``psi``: (dim,) or (batch, dim) complex -> (D, D) or (batch, D, D)."""
        psi = _c128(psi)
        single = psi.ndim == 1
        psi2 = psi.reshape(1, -1) if single else psi
        if psi2.ndim != 2:
            raise ValueError("psi must be 1-D or 2-D")
        batch, dim = psi2.shape
        n = _n_qubits(dim)
        traced = list(traced)
        D = 1 << (n - len(traced)) if len(traced) <= n else 0
        out = np.empty((batch, max(D, 0), max(D, 0)), dtype=np.complex128)
        check(self._lib.bq_partial_trace_sv_host(self._h, _ptr(psi2), n, _int_array(traced), len(traced), batch, dim,
                                                 _ptr(out)))
        return out[0] if single else out

    def partial_trace_sv_multi(self, psi, traced_sets: Sequence[Sequence[int]]) -> np.ndarray:
        """Several traces of one statevector (all sets the same length) -> (n_sets, D, D)."""
        psi = _c128(psi).reshape(-1)
        n = _n_qubits(psi.shape[0])
        sets = [list(s) for s in traced_sets]
        if not sets:
            return np.empty((0, 0, 0), dtype=np.complex128)
        nt = len(sets[0])
        if any(len(s) != nt for s in sets):
            raise ValueError("all traced sets must have the same length")
        D = 1 << (n - nt) if nt <= n else 0
        out = np.empty((len(sets), D, D), dtype=np.complex128)
        flat = [q for s in sets for q in s]
        check(self._lib.bq_partial_trace_sv_multi_host(self._h, _ptr(psi), n, _int_array(flat), nt, len(sets),
                                                       _ptr(out)))
        return out

    def partial_trace_sv_mean(self, psi, traced: Sequence[int], weights=None) -> np.ndarray:
        """Shot-averaged reduced density matrix: ``psi`` (shots, dim) -> ``sum_k w_k Tr_traced |psi_k><psi_k|`` (D, D),
        ``weights=None`` being the plain mean.  One batched launch (``bq_partial_trace_sv_accumulate_host``)."""
        psi = _c128(psi)
        if psi.ndim != 2 or psi.shape[0] == 0:
            raise ValueError("psi must be (shots, dim) with at least one shot")
        shots, dim = psi.shape
        n = _n_qubits(dim)
        traced = list(traced)
        D = 1 << (n - len(traced)) if len(traced) <= n else 0
        w = None
        if weights is not None:
            w = _f64(weights).reshape(-1)
            if w.shape[0] != shots:
                raise ValueError("weights must have one entry per shot")
        out = np.empty((D, D), dtype=np.complex128)
        check(self._lib.bq_partial_trace_sv_accumulate_host(self._h, _ptr(psi), n, _int_array(traced), len(traced), shots,
                                                            dim, _ptr(w) if w is not None else None, _ptr(out)))
        return out

    def partial_trace_sv_mean_t(self, psi, traced: Sequence[int], weights=None, out=None, stream=None):
        """Device-resident ``partial_trace_sv_mean``: ``psi`` (shots, dim) complex128 CUDA tensor -> (D, D) tensor."""
        import torch

        self._check_tensor(psi, torch.complex128, "psi")
        if psi.dim() != 2 or psi.shape[0] == 0:
            raise ValueError("psi must be (shots, dim) with at least one shot")
        shots, dim = psi.shape
        n = _n_qubits(dim)
        traced = list(traced)
        D = 1 << max(n - len(traced), 0)
        if weights is not None:
            self._check_tensor(weights, torch.float64, "weights")
        if out is None:
            out = torch.empty((D, D), dtype=torch.complex128, device=psi.device)
        check(self._lib.bq_partial_trace_sv_accumulate_dev(
            self._h, ctypes.c_void_p(psi.data_ptr()), n, _int_array(traced), len(traced), shots, dim,
            ctypes.c_void_p(weights.data_ptr()) if weights is not None else None, ctypes.c_void_p(out.data_ptr()),
            self._stream_ptr(stream)))
        return out

    def state_to_wigner_multi(self, psi, traced_sets: Sequence[Sequence[int]], xvec, yvec=None, g: float = math.sqrt(2),
                              method: str = "clenshaw", return_rho: bool = False, out=None):
        """Several traces of ONE statevector, each followed by its Wigner grid, in one call (per-site reduced density
        matrices + per-site Wigner functions): ``psi`` (dim,) -> W (n_sets, ny, nx) [, rho (n_sets, D, D)]."""
        code = method_code(method)
        psi = _c128(psi).reshape(-1)
        n = _n_qubits(psi.shape[0])
        sets = [list(s) for s in traced_sets]
        nt = len(sets[0]) if sets else 0
        if any(len(s) != nt for s in sets):
            raise ValueError("all traced sets must have the same length")
        if code == METHODS["fft"]:
            rho_f = self.partial_trace_sv_multi(psi, sets)
            res = self.wigner_fft(rho_f, xvec, g=g)
            return (res, rho_f) if return_rho else res
        xvec = _f64(xvec).reshape(-1)
        yvec = xvec if yvec is None else _f64(yvec).reshape(-1)
        shape = (len(sets), len(yvec), len(xvec))
        if out is None:
            out = self.pinned_empty(shape, np.float64)
        elif out.shape != shape or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("out has the wrong shape / dtype / layout")
        D = 1 << max(n - nt, 0)
        rho = np.empty((len(sets), D, D), dtype=np.complex128) if return_rho else None
        flat = [q for s in sets for q in s]
        check(self._lib.bq_state_to_wigner_multi_host(
            self._h, _ptr(psi), n, _int_array(flat), nt, len(sets), _ptr(xvec), len(xvec), _ptr(yvec), len(yvec), float(g),
            code, _ptr(out), _ptr(rho) if return_rho else None))
        return (out, rho) if return_rho else out

    def partial_trace_dm(self, rho, traced: Sequence[int]) -> np.ndarray:
        rho = _c128(rho)
        single = rho.ndim == 2
        rho3 = rho.reshape(1, *rho.shape) if single else rho
        if rho3.ndim != 3 or rho3.shape[1] != rho3.shape[2]:
            raise ValueError("rho must be (F, F) or (batch, F, F)")
        batch, F, _ = rho3.shape
        n = _n_qubits(F)
        traced = list(traced)
        D = 1 << (n - len(traced)) if len(traced) <= n else 0
        out = np.empty((batch, D, D), dtype=np.complex128)
        check(self._lib.bq_partial_trace_dm_host(self._h, _ptr(rho3), n, _int_array(traced), len(traced), batch,
                                                 _ptr(out)))
        return out[0] if single else out

    def photon_number(self, rho) -> np.ndarray:
        """Unrounded complex <n> per density matrix: (D, D) -> scalar, (batch, D, D) -> (batch,)."""
        rho = _c128(rho)
        single = rho.ndim == 2
        rho3 = rho.reshape(1, *rho.shape) if single else rho
        batch, D, _ = rho3.shape
        out = np.empty((batch,), dtype=np.complex128)
        check(self._lib.bq_photon_number_host(self._h, _ptr(rho3), D, batch, _ptr(out)))
        return out[0] if single else out

    # ------------------------------------------------------------------ Wigner (host)
    def wigner(self, rho, xvec, yvec=None, g: float = math.sqrt(2), method: str = "clenshaw", out=None) -> np.ndarray:
        """``rho``: (N, N) or (batch, N, N) -> W (ny, nx) or (batch, ny, nx), indexed [iy, ix]."""
        code = method_code(method)
        if code == METHODS["fft"]:  # returns the tuple (W, yvec) on its own p axis, like qutip.wigner
            return self.wigner_fft(rho, xvec, g=g)
        rho = _c128(rho)
        single = rho.ndim == 2
        rho3 = rho.reshape(1, *rho.shape) if single else rho
        if rho3.ndim != 3 or rho3.shape[1] != rho3.shape[2]:
            raise TypeError("Input state is not a valid operator.")
        batch, N, _ = rho3.shape
        xvec = _f64(xvec).reshape(-1)
        yvec = xvec if yvec is None else _f64(yvec).reshape(-1)
        shape = (batch, len(yvec), len(xvec))
        if out is None:
            out = self.pinned_empty(shape, np.float64)  # the kernel writes host memory directly
        elif out.shape != shape or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("out has the wrong shape / dtype / layout")
        check(self._lib.bq_wigner_host(self._h, _ptr(rho3), N, N, N * N, batch, _ptr(xvec), len(xvec), _ptr(yvec),
                                       len(yvec), float(g), code, _ptr(out)))
        return out[0] if single else out

    def wigner_fft(self, rho, xvec, g: float = math.sqrt(2)):
        """``qutip.wigner(..., method="fft")``: ``rho`` (N, N) or (batch, N, N) -> ``(W, yvec)`` with ``W[ip, ix]`` of
        shape (S, S) or (batch, S, S) on the p axis ``yvec`` the method derives from the (uniform) x grid."""
        rho = _c128(rho)
        single = rho.ndim == 2
        rho3 = rho.reshape(1, *rho.shape) if single else rho
        if rho3.ndim != 3 or rho3.shape[1] != rho3.shape[2]:
            raise TypeError("Input state is not a valid operator.")
        batch, N, _ = rho3.shape
        xvec = _f64(xvec).reshape(-1)
        S = len(xvec)
        out = np.empty((batch, S, S), dtype=np.float64)
        yvec = np.empty((S,), dtype=np.float64)
        check(self._lib.bq_wigner_fft_host(self._h, _ptr(rho3), N, N, N * N, batch, _ptr(xvec), S, float(g), _ptr(out),
                                           _ptr(yvec)))
        return (out[0] if single else out), yvec

    def wigner_fft_t(self, rho, xvec, g: float = math.sqrt(2), out=None, stream=None):
        """Device-resident ``method="fft"``: ``rho`` complex128 CUDA tensor, ``xvec`` a HOST array (the p axis is
        computed from it on the host) -> ``(W tensor, yvec ndarray)``."""
        import torch

        self._check_tensor(rho, torch.complex128, "rho")
        single = rho.dim() == 2
        rho3 = rho.unsqueeze(0) if single else rho
        batch, N, _ = rho3.shape
        xvec = _f64(xvec).reshape(-1)
        S = len(xvec)
        if out is None:
            out = torch.empty((batch, S, S), dtype=torch.float64, device=rho.device)
        else:
            self._check_tensor(out, torch.float64, "out")
        yvec = np.empty((S,), dtype=np.float64)
        check(self._lib.bq_wigner_fft_dev(self._h, ctypes.c_void_p(rho3.data_ptr()), N, N, N * N, batch, _ptr(xvec), S,
                                          float(g), ctypes.c_void_p(out.data_ptr()), _ptr(yvec), self._stream_ptr(stream)))
        return (out[0] if single else out), yvec

    def state_to_wigner(self, psi, traced: Sequence[int], xvec, yvec=None, g: float = math.sqrt(2),
                        method: str = "clenshaw", return_rho: bool = False, out=None):
        """Fused frame loop: ``psi`` (dim,) or (frames, dim) -> W (and optionally the reduced rho).
        ``method="fft"`` returns ``(W, yvec)`` in place of ``W`` (trace, then ``wigner_fft``)."""
        code = method_code(method)
        if code == METHODS["fft"]:
            rho_f = self.partial_trace_sv(psi, traced)
            res = self.wigner_fft(rho_f, xvec, g=g)
            return (res, rho_f) if return_rho else res
        psi = _c128(psi)
        single = psi.ndim == 1
        psi2 = psi.reshape(1, -1) if single else psi
        batch, dim = psi2.shape
        n = _n_qubits(dim)
        traced = list(traced)
        xvec = _f64(xvec).reshape(-1)
        yvec = xvec if yvec is None else _f64(yvec).reshape(-1)
        shape = (batch, len(yvec), len(xvec))
        if out is None:
            out = self.pinned_empty(shape, np.float64)  # the kernel writes host memory directly
        elif out.shape != shape or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("out has the wrong shape / dtype / layout")
        D = 1 << max(n - len(traced), 0)
        rho = np.empty((batch, D, D), dtype=np.complex128) if return_rho else None
        check(self._lib.bq_state_to_wigner_host(
            self._h, _ptr(psi2), n, _int_array(traced), len(traced), batch, dim, _ptr(xvec), len(xvec), _ptr(yvec),
            len(yvec), float(g), code, _ptr(out), _ptr(rho) if return_rho else None))
        W = out[0] if single else out
        if return_rho:
            return W, (rho[0] if single else rho)
        return W

    # ------------------------------------------------------------------ device-resident (torch tensors)
    @staticmethod
    def _stream_ptr(stream):
        import torch

        s = torch.cuda.current_stream() if stream is None else stream
        return ctypes.c_void_p(s.cuda_stream)

    def _check_tensor(self, t, dtype, name):
        import torch

        if not isinstance(t, torch.Tensor) or not t.is_cuda or t.device.index != self.device:
            raise ValueError(f"{name} must be a CUDA tensor on device {self.device}")
        if t.dtype != dtype or not t.is_contiguous():
            raise ValueError(f"{name} must be contiguous {dtype}")

    def grid_units(self, S: int) -> int:
        """Units (pairs of grid lines, 8 points each) of an S x S mirror-symmetric grid."""
        return int(self._lib.bq_grid_units(int(S)))

    def wigner_t(self, rho, xvec, yvec=None, g: float = math.sqrt(2), method: str = "clenshaw", out=None, stream=None,
                 unit_range=None):
        """Device-resident Wigner: rho (batch, N, N) complex128 CUDA tensor -> (batch, ny, nx) float64 tensor.

        ``unit_range=(first, count)``: evaluate only that share of a mirror-symmetric grid's units (multi-GPU sharding
        of one large grid); the other points of ``out`` are left untouched.  Raises ``BQError`` (``BQ_ERR_UNSUPPORTED``)
        when the grid or cutoff cannot use the symmetric-grid kernels."""
        import torch

        code = method_code(method)
        self._check_tensor(rho, torch.complex128, "rho")
        single = rho.dim() == 2
        rho3 = rho.unsqueeze(0) if single else rho
        batch, N, _ = rho3.shape
        yvec = xvec if yvec is None else yvec
        self._check_tensor(xvec, torch.float64, "xvec")
        self._check_tensor(yvec, torch.float64, "yvec")
        if out is None:
            out = torch.empty((batch, yvec.numel(), xvec.numel()), dtype=torch.float64, device=rho.device)
        else:
            self._check_tensor(out, torch.float64, "out")
        with self._grid_lock:
            check(self._lib.bq_set_grid_symmetry(self._h, self._device_grid_mode(xvec, yvec)))
            if unit_range is not None:
                check(self._lib.bq_set_unit_range(self._h, int(unit_range[0]), int(unit_range[1])))
            try:
                check(self._lib.bq_wigner_dev(self._h, ctypes.c_void_p(rho3.data_ptr()), N, N, N * N, batch,
                                              ctypes.c_void_p(xvec.data_ptr()), xvec.numel(),
                                              ctypes.c_void_p(yvec.data_ptr()), yvec.numel(), float(g), code,
                                              ctypes.c_void_p(out.data_ptr()), self._stream_ptr(stream)))
            finally:
                if unit_range is not None:
                    check(self._lib.bq_set_unit_range(self._h, 0, -1))
        return out[0] if single else out

    def partial_trace_sv_t(self, psi, traced: Sequence[int], out=None, stream=None):
        import torch

        self._check_tensor(psi, torch.complex128, "psi")
        single = psi.dim() == 1
        psi2 = psi.unsqueeze(0) if single else psi
        batch, dim = psi2.shape
        n = _n_qubits(dim)
        traced = list(traced)
        D = 1 << max(n - len(traced), 0)
        if out is None:
            out = torch.empty((batch, D, D), dtype=torch.complex128, device=psi.device)
        check(self._lib.bq_partial_trace_sv_dev(self._h, ctypes.c_void_p(psi2.data_ptr()), n, _int_array(traced),
                                                len(traced), batch, dim, ctypes.c_void_p(out.data_ptr()),
                                                self._stream_ptr(stream)))
        return out[0] if single else out

    def state_to_wigner_t(self, psi, traced: Sequence[int], xvec, yvec=None, g: float = math.sqrt(2),
                          method: str = "clenshaw", out=None, rho_out=None, stream=None):
        """Device-resident fused frame loop; ``out`` may alias peer memory opened with ``ipc_open``."""
        import torch

        code = method_code(method)
        self._check_tensor(psi, torch.complex128, "psi")
        single = psi.dim() == 1
        psi2 = psi.unsqueeze(0) if single else psi
        batch, dim = psi2.shape
        n = _n_qubits(dim)
        traced = list(traced)
        yvec = xvec if yvec is None else yvec
        self._check_tensor(xvec, torch.float64, "xvec")
        self._check_tensor(yvec, torch.float64, "yvec")
        out_ptr = None
        stage = None
        if out is None:
            out = torch.empty((batch, yvec.numel(), xvec.numel()), dtype=torch.float64, device=psi.device)
            out_ptr = out.data_ptr()
        elif isinstance(out, int):
            out_ptr = out  # raw device pointer (e.g. a peer buffer)
        else:
            out_ptr = out.data_ptr()
        with self._grid_lock:
            mode = self._device_grid_mode(xvec, yvec)
            if isinstance(out, int) and mode == 2:
                # A raw output pointer is a peer GPU's buffer.  The point-by-point kernels store finished tiles into
                # it directly (contiguous bulk stores, fused compute + gather); the symmetric-grid kernels scatter
                # 8-byte stores, so they fill a local block that then crosses NVLink as one copy on the same stream.
                n_out = batch * yvec.numel() * xvec.numel()
                if self._stage is None or self._stage.numel() < n_out or self._stage.device != psi.device:
                    self._stage = torch.empty((n_out,), dtype=torch.float64, device=psi.device)
                stage = self._stage
                if stream is not None:
                    # the block is read by a copy on `stream`: tell torch's caching allocator, so that a later
                    # re-allocation of the staging block cannot hand its memory out while that copy is in flight
                    stage.record_stream(stream)
                out_ptr = stage.data_ptr()
            check(self._lib.bq_set_grid_symmetry(self._h, mode))
            check(self._lib.bq_state_to_wigner_dev(
                self._h, ctypes.c_void_p(psi2.data_ptr()), n, _int_array(traced), len(traced), batch, dim,
                ctypes.c_void_p(xvec.data_ptr()), xvec.numel(), ctypes.c_void_p(yvec.data_ptr()), yvec.numel(), float(g),
                code, ctypes.c_void_p(out_ptr), ctypes.c_void_p(rho_out.data_ptr()) if rho_out is not None else None,
                self._stream_ptr(stream)))
            if stage is not None:
                check(self._lib.bq_memcpy_dev(self._h, ctypes.c_void_p(out), ctypes.c_void_p(out_ptr),
                                              batch * yvec.numel() * xvec.numel() * 8, self._stream_ptr(stream)))
        if isinstance(out, int):
            return None
        return out[0] if single else out

    def state_to_wigner_multi_t(self, psi, traced_sets: Sequence[Sequence[int]], xvec, yvec=None, g: float = math.sqrt(2),
                                method: str = "clenshaw", out=None, rho_out=None, stream=None):
        """Device-resident ``state_to_wigner_multi``: one statevector tensor -> (n_sets, ny, nx) tensor."""
        import torch

        code = method_code(method)
        self._check_tensor(psi, torch.complex128, "psi")
        psi = psi.reshape(-1)
        n = _n_qubits(psi.numel())
        sets = [list(s) for s in traced_sets]
        nt = len(sets[0]) if sets else 0
        if any(len(s) != nt for s in sets):
            raise ValueError("all traced sets must have the same length")
        yvec = xvec if yvec is None else yvec
        self._check_tensor(xvec, torch.float64, "xvec")
        self._check_tensor(yvec, torch.float64, "yvec")
        if out is None:
            out = torch.empty((len(sets), yvec.numel(), xvec.numel()), dtype=torch.float64, device=psi.device)
        flat = [q for s in sets for q in s]
        with self._grid_lock:
            check(self._lib.bq_set_grid_symmetry(self._h, self._device_grid_mode(xvec, yvec)))
            check(self._lib.bq_state_to_wigner_multi_dev(
                self._h, ctypes.c_void_p(psi.data_ptr()), n, _int_array(flat), nt, len(sets), ctypes.c_void_p(xvec.data_ptr()),
                xvec.numel(), ctypes.c_void_p(yvec.data_ptr()), yvec.numel(), float(g), code, ctypes.c_void_p(out.data_ptr()),
                ctypes.c_void_p(rho_out.data_ptr()) if rho_out is not None else None, self._stream_ptr(stream)))
        return out

    def frame_minmax_t(self, W, stream=None):
        import torch

        self._check_tensor(W, torch.float64, "W")
        batch = W.shape[0]
        points = W[0].numel() if batch else 0
        out = torch.empty((batch, 2), dtype=torch.float64, device=W.device)
        check(self._lib.bq_frame_minmax_dev(self._h, ctypes.c_void_p(W.data_ptr()), points, batch,
                                            ctypes.c_void_p(out.data_ptr()), self._stream_ptr(stream)))
        return out

    def frame_rgba_t(self, W, levels: int = 100, zero_abs_max: float = 5.0, out=None, stream=None):
        """(frames, ny, nx) float64 CUDA tensor -> (frames, ny, nx, 4) uint8: the RdBu colour ``contourf`` gives each
        grid node with the reference's symmetric colour levels (``animate.py:278-291``)."""
        import torch

        self._check_tensor(W, torch.float64, "W")
        W3 = W.unsqueeze(0) if W.dim() == 2 else W
        batch = W3.shape[0]
        points = W3[0].numel() if batch else 0
        if out is None:
            out = torch.empty((*W3.shape, 4), dtype=torch.uint8, device=W.device)
        check(self._lib.bq_frame_rgba_dev(self._h, ctypes.c_void_p(W3.data_ptr()), points, batch, int(levels),
                                          float(zero_abs_max), ctypes.c_void_p(out.data_ptr()), self._stream_ptr(stream)))
        return out[0] if W.dim() == 2 else out

    # ------------------------------------------------------------------ CUDA IPC
    def device_alloc(self, nbytes: int) -> int:
        """Raw cudaMalloc on this engine's device (whole allocation: exportable through CUDA IPC)."""
        p = ctypes.c_void_p()
        check(self._lib.bq_device_alloc(self._h, int(nbytes), ctypes.byref(p)))
        return int(p.value)

    def device_free(self, ptr: int) -> None:
        check(self._lib.bq_device_free(self._h, ctypes.c_void_p(ptr)))

    def ipc_export(self, ptr) -> bytes:
        """64-byte CUDA IPC handle of an allocation made with ``device_alloc`` (or a tensor's storage base)."""
        if not isinstance(ptr, int):
            ptr = ptr.data_ptr()
        buf = (ctypes.c_ubyte * 64)()
        check(self._lib.bq_ipc_export(self._h, ctypes.c_void_p(ptr), buf))
        return bytes(buf)

    def ipc_open(self, handle: bytes) -> int:
        buf = (ctypes.c_ubyte * 64).from_buffer_copy(handle)
        p = ctypes.c_void_p()
        check(self._lib.bq_ipc_open(self._h, buf, ctypes.byref(p)))
        return int(p.value)

    def ipc_close(self, ptr: int) -> None:
        check(self._lib.bq_ipc_close(self._h, ctypes.c_void_p(ptr)))


class MultiEngine:
    """All (or some of) the GPUs of one box from ONE process: owns a ``bq_multi_t`` (one handle + one worker thread per
    device).  Replaces the reference's ``multiprocessing.Pool`` over frames (``animate.py:166-169,191-206``).

    ``shard="frames"``: contiguous frame blocks per GPU, each GPU writes its finished grids into its slice of the
    (page-locked) result; ``shard="grid"``: ONE large grid shared out over the GPUs (``BQ_SHARD_GRID``)."""

    def __init__(self, devices: Sequence[int] | None = None):
        self._lib = _lib.load()
        m = ctypes.c_void_p()
        if devices is None:
            check(self._lib.bq_multi_init(ctypes.byref(m), 0, None))
        else:
            devs = [int(d) for d in devices]
            check(self._lib.bq_multi_init(ctypes.byref(m), len(devs), _int_array(devs)))
        self._m = m
        self._finalizer = weakref.finalize(self, self._lib.bq_multi_finalize, m)
        n = ctypes.c_int()
        check(self._lib.bq_multi_device_count(self._m, ctypes.byref(n)))
        self.n_devices = int(n.value)
        h0 = ctypes.c_void_p()
        check(self._lib.bq_multi_handle(self._m, 0, ctypes.byref(h0)))
        self._h0 = h0  # page-locked buffers are portable: allocated through the first device's handle
        self._pinned: dict[tuple, np.ndarray] = {}

    def close(self) -> None:
        self._pinned.clear()
        self._finalizer()

    @property
    def launches(self) -> int:
        n = ctypes.c_uint64()
        check(self._lib.bq_multi_launch_count(self._m, ctypes.byref(n)))
        return int(n.value)

    def handle(self, index: int):
        h = ctypes.c_void_p()
        check(self._lib.bq_multi_handle(self._m, int(index), ctypes.byref(h)))
        return h

    def set_sass_mode(self, patched: bool) -> None:
        for k in range(self.n_devices):
            check(self._lib.bq_set_sass_mode(self.handle(k), 1 if patched else 0, None))

    def pinned_empty(self, shape, dtype=np.float64) -> np.ndarray:
        """Page-locked host array (portable across the devices); freed when garbage collected."""
        dtype = np.dtype(dtype)
        count = int(np.prod(shape, dtype=np.int64))
        nbytes = max(count * dtype.itemsize, 1)
        p = ctypes.c_void_p()
        check(self._lib.bq_host_alloc(self._h0, nbytes, ctypes.byref(p)))
        buf = (ctypes.c_char * nbytes).from_address(int(p.value))
        arr = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
        weakref.finalize(buf, MultiEngine._free_pinned, weakref.ref(self), int(p.value))
        return arr

    @staticmethod
    def _free_pinned(self_ref, ptr: int) -> None:
        self = self_ref()
        if self is None or not self._finalizer.alive:
            return
        self._lib.bq_host_free(self._h0, ctypes.c_void_p(ptr))

    @staticmethod
    def _mode(shard: str) -> int:
        try:
            return {"frames": _lib.BQ_SHARD_FRAMES, "grid": _lib.BQ_SHARD_GRID}[shard]
        except KeyError:
            raise ValueError("shard must be 'frames' or 'grid'") from None

    def state_to_wigner(self, psi, traced: Sequence[int], xvec, yvec=None, g: float = math.sqrt(2),
                        method: str = "clenshaw", shard: str = "frames", return_rho: bool = False, out=None):
        """``Engine.state_to_wigner`` over all GPUs: ``psi`` (dim,) or (frames, dim) -> W [, rho]."""
        code = method_code(method)
        if code == METHODS["fft"]:
            raise BQError(_lib.BQ_ERR_UNSUPPORTED, "method 'fft' is not sharded: use Engine.state_to_wigner")
        psi = _c128(psi)
        single = psi.ndim == 1
        psi2 = psi.reshape(1, -1) if single else psi
        batch, dim = psi2.shape
        n = _n_qubits(dim)
        traced = list(traced)
        xvec = _f64(xvec).reshape(-1)
        yvec = xvec if yvec is None else _f64(yvec).reshape(-1)
        shape = (batch, len(yvec), len(xvec))
        if out is None:
            out = self.pinned_empty(shape, np.float64)
        elif out.shape != shape or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("out has the wrong shape / dtype / layout")
        D = 1 << max(n - len(traced), 0)
        rho = np.empty((batch, D, D), dtype=np.complex128) if return_rho else None
        check(self._lib.bq_state_to_wigner_sharded_host(
            self._m, _ptr(psi2), n, _int_array(traced), len(traced), batch, dim, _ptr(xvec), len(xvec), _ptr(yvec), len(yvec),
            float(g), code, self._mode(shard), _ptr(out), _ptr(rho) if return_rho else None))
        W = out[0] if single else out
        return (W, (rho[0] if single else rho)) if return_rho else W

    def wigner(self, rho, xvec, yvec=None, g: float = math.sqrt(2), method: str = "clenshaw", shard: str = "frames", out=None):
        """``Engine.wigner`` over all GPUs: ``rho`` (N, N) or (batch, N, N) -> W."""
        code = method_code(method)
        if code == METHODS["fft"]:
            raise BQError(_lib.BQ_ERR_UNSUPPORTED, "method 'fft' is not sharded: use Engine.wigner")
        rho = _c128(rho)
        single = rho.ndim == 2
        rho3 = rho.reshape(1, *rho.shape) if single else rho
        if rho3.ndim != 3 or rho3.shape[1] != rho3.shape[2]:
            raise TypeError("Input state is not a valid operator.")
        batch, N, _ = rho3.shape
        xvec = _f64(xvec).reshape(-1)
        yvec = xvec if yvec is None else _f64(yvec).reshape(-1)
        shape = (batch, len(yvec), len(xvec))
        if out is None:
            out = self.pinned_empty(shape, np.float64)
        elif out.shape != shape or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("out has the wrong shape / dtype / layout")
        check(self._lib.bq_wigner_sharded_host(self._m, _ptr(rho3), N, N, N * N, batch, _ptr(xvec), len(xvec), _ptr(yvec),
                                               len(yvec), float(g), code, self._mode(shard), _ptr(out)))
        return out[0] if single else out


_multi_engines: dict[tuple, MultiEngine] = {}


def get_multi_engine(devices: Sequence[int] | None = None) -> MultiEngine:
    """Process-wide ``MultiEngine`` over ``devices`` (None: every visible GPU).  Raises without a GPU."""
    key = tuple(int(d) for d in devices) if devices is not None else ("all",)
    with _engines_lock:
        me = _multi_engines.get(key)
        if me is None:
            me = MultiEngine(devices)
            _multi_engines[key] = me
        return me


def device_count() -> int:
    """CUDA devices the library sees (0 without a driver / GPU)."""
    n = ctypes.c_int()
    rc = _lib.load().bq_device_count(ctypes.byref(n))
    return int(n.value) if rc == 0 else 0


_engines: dict[int, Engine] = {}
_engines_lock = threading.Lock()


def get_engine(device: int | None = None) -> Engine:
    """Process-wide engine for ``device`` (default: LOCAL_RANK if set, else 0).  Raises without a GPU."""
    import os

    if device is None:
        device = int(os.environ.get("BQ_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    with _engines_lock:
        eng = _engines.get(device)
        if eng is None:
            eng = Engine(device)
            _engines[device] = eng
        return eng


__all__ = ["Engine", "MultiEngine", "get_engine", "get_multi_engine", "device_count", "BQError", "method_code", "METHODS"]
