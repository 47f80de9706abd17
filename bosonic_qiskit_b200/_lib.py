"""ctypes binding of ``libbq_b200.so`` (C ABI declared in ``include/bq_b200.h``).

The product path has no CPU fallback: if the library is missing, or no B200 is
visible, the calls raise instead of computing on the host.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int64, c_size_t, c_uint64, c_void_p

from ._build import LIB_PATH

_dp = POINTER(c_double)
_ip = POINTER(c_int)
_H = c_void_p

# name -> (restype, argtypes); mirrors include/bq_b200.h one to one
SIGNATURES = {
    "bq_version": (c_int, []),
    "bq_last_error": (c_char_p, []),
    "bq_device_count": (c_int, [_ip]),
    "bq_create": (c_int, [POINTER(_H), c_int]),
    "bq_destroy": (c_int, [_H]),
    "bq_device": (c_int, [_H, _ip]),
    "bq_launch_count": (c_int, [_H, POINTER(c_uint64)]),
    "bq_synchronize": (c_int, [_H]),
    "bq_host_alloc": (c_int, [_H, c_size_t, POINTER(c_void_p)]),
    "bq_host_free": (c_int, [_H, c_void_p]),
    "bq_partial_trace_sv_host": (c_int, [_H, c_void_p, c_int, _ip, c_int, c_int, c_int64, c_void_p]),
    "bq_partial_trace_sv_dev": (c_int, [_H, c_void_p, c_int, _ip, c_int, c_int, c_int64, c_void_p, c_void_p]),
    "bq_partial_trace_sv_multi_host": (c_int, [_H, c_void_p, c_int, _ip, c_int, c_int, c_void_p]),
    "bq_partial_trace_sv_multi_dev": (c_int, [_H, c_void_p, c_int, _ip, c_int, c_int, c_void_p, c_void_p]),
    "bq_partial_trace_sv_accumulate_host": (c_int, [_H, c_void_p, c_int, _ip, c_int, c_int, c_int64, c_void_p, c_void_p]),
    "bq_partial_trace_sv_accumulate_dev": (c_int, [_H, c_void_p, c_int, _ip, c_int, c_int, c_int64, c_void_p, c_void_p,
                                                   c_void_p]),
    "bq_partial_trace_dm_host": (c_int, [_H, c_void_p, c_int, _ip, c_int, c_int, c_void_p]),
    "bq_partial_trace_dm_dev": (c_int, [_H, c_void_p, c_int, _ip, c_int, c_int, c_void_p, c_void_p]),
    "bq_photon_number_host": (c_int, [_H, c_void_p, c_int, c_int, c_void_p]),
    "bq_photon_number_dev": (c_int, [_H, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "bq_wigner_host": (c_int, [_H, c_void_p, c_int, c_int64, c_int64, c_int, c_void_p, c_int, c_void_p, c_int,
                               c_double, c_int, c_void_p]),
    "bq_wigner_dev": (c_int, [_H, c_void_p, c_int, c_int64, c_int64, c_int, c_void_p, c_int, c_void_p, c_int,
                              c_double, c_int, c_void_p, c_void_p]),
    "bq_wigner_fft_host": (c_int, [_H, c_void_p, c_int, c_int64, c_int64, c_int, c_void_p, c_int, c_double, c_void_p,
                                   c_void_p]),
    "bq_wigner_fft_dev": (c_int, [_H, c_void_p, c_int, c_int64, c_int64, c_int, c_void_p, c_int, c_double, c_void_p,
                                  c_void_p, c_void_p]),
    "bq_state_to_wigner_host": (c_int, [_H, c_void_p, c_int, _ip, c_int, c_int, c_int64, c_void_p, c_int, c_void_p,
                                        c_int, c_double, c_int, c_void_p, c_void_p]),
    "bq_state_to_wigner_dev": (c_int, [_H, c_void_p, c_int, _ip, c_int, c_int, c_int64, c_void_p, c_int, c_void_p,
                                       c_int, c_double, c_int, c_void_p, c_void_p, c_void_p]),
    "bq_state_to_wigner_multi_host": (c_int, [_H, c_void_p, c_int, _ip, c_int, c_int, c_void_p, c_int, c_void_p, c_int,
                                              c_double, c_int, c_void_p, c_void_p]),
    "bq_state_to_wigner_multi_dev": (c_int, [_H, c_void_p, c_int, _ip, c_int, c_int, c_void_p, c_int, c_void_p, c_int,
                                             c_double, c_int, c_void_p, c_void_p, c_void_p]),
    "bq_frame_minmax_dev": (c_int, [_H, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "bq_frame_rgba_dev": (c_int, [_H, c_void_p, c_int64, c_int, c_int, c_double, c_void_p, c_void_p]),
    "bq_rdbu_lut": (c_int, [c_void_p]),
    "bq_fp64_peak": (c_int, [_H, c_double, _dp]),
    "bq_last_device_ms": (c_int, [_H, _dp]),
    "bq_profile_enable": (c_int, [_H, c_int]),
    "bq_set_sass_mode": (c_int, [_H, c_int, _ip]),
    "bq_sass_self_check": (c_int, [_H, POINTER(c_uint64), _ip]),
    "bq_profile_read": (c_int, [_H, _dp, _ip]),
    "bq_set_grid_symmetry": (c_int, [_H, c_int]),
    "bq_set_unit_range": (c_int, [_H, c_int64, c_int64]),
    "bq_grid_units": (c_int64, [c_int]),
    "bq_set_sym_variant": (c_int, [_H, c_int]),
    "bq_last_wigner_variant": (c_int, [_H, _ip]),
    "bq_grid_is_mirror": (c_int, [c_void_p, c_int, c_void_p, c_int]),
    "bq_multi_init": (c_int, [POINTER(_H), c_int, _ip]),
    "bq_multi_finalize": (c_int, [_H]),
    "bq_multi_device_count": (c_int, [_H, _ip]),
    "bq_multi_handle": (c_int, [_H, c_int, POINTER(_H)]),
    "bq_multi_launch_count": (c_int, [_H, POINTER(c_uint64)]),
    "bq_state_to_wigner_sharded_host": (c_int, [_H, c_void_p, c_int, _ip, c_int, c_int, c_int64, c_void_p, c_int, c_void_p,
                                                c_int, c_double, c_int, c_int, c_void_p, c_void_p]),
    "bq_wigner_sharded_host": (c_int, [_H, c_void_p, c_int, c_int64, c_int64, c_int, c_void_p, c_int, c_void_p, c_int,
                                       c_double, c_int, c_int, c_void_p]),
    "bq_device_alloc": (c_int, [_H, c_size_t, POINTER(c_void_p)]),
    "bq_device_free": (c_int, [_H, c_void_p]),
    "bq_ipc_export": (c_int, [_H, c_void_p, c_void_p]),
    "bq_ipc_open": (c_int, [_H, c_void_p, POINTER(c_void_p)]),
    "bq_ipc_close": (c_int, [_H, c_void_p]),
    "bq_memcpy_dev": (c_int, [_H, c_void_p, c_void_p, c_size_t, c_void_p]),
}

BQ_OK, BQ_ERR_INVALID, BQ_ERR_CUDA, BQ_ERR_NO_DEVICE, BQ_ERR_UNSUPPORTED = 0, -1, -2, -3, -4
BQ_SHARD_FRAMES, BQ_SHARD_GRID = 0, 1

_lib = None


class BQError(RuntimeError):
    """A call into libbq_b200 failed; ``code`` is the negative ``bq_status``."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libbq_b200 error {code}: {message}")
        self.code = code
        self.message = message


def library_path() -> str:
    return os.environ.get("BQ_B200_LIBRARY", LIB_PATH)


def load():
    """Load the shared library and declare every prototype of include/bq_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -m bosonic_qiskit_b200._build` "
            "(nvcc, sm_100a).  There is no CPU fallback for this path."
        )
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != BQ_OK:
        msg = load().bq_last_error()
        raise BQError(rc, msg.decode("utf-8", "replace") if msg else "unknown error")
