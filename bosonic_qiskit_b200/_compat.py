"""Optional third-party types.

qiskit is the reference's container for states (``Statevector`` / ``DensityMatrix``).  When it is
importable the drop-in functions return the real classes so downstream code is untouched; when it
is not (this build image has no qiskit) minimal stand-ins with the same attributes the hot path
touches are used, so the package stays importable and testable.
"""

from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the environment
    from qiskit.quantum_info import DensityMatrix, Statevector  # type: ignore

    HAVE_QISKIT = True
except Exception:  # ModuleNotFoundError here
    HAVE_QISKIT = False

    def _qubit_dims(dim: int) -> tuple[int, ...]:
        n = int(dim).bit_length() - 1
        return (2,) * n if (1 << n) == dim else (dim,)

    class Statevector:  # noqa: D401 - stand-in
        """Stand-in for ``qiskit.quantum_info.Statevector`` (only what the hot path uses)."""

        def __init__(self, data, dims=None):
            self._data = np.asarray(getattr(data, "data", data), dtype=np.complex128).reshape(-1)
            self._dims = tuple(dims) if dims is not None else _qubit_dims(self._data.shape[0])

        @property
        def data(self):
            return self._data

        @property
        def dim(self):
            return self._data.shape[0]

        def dims(self):
            return self._dims

        def inner(self, other):
            return np.vdot(self._data, np.asarray(getattr(other, "data", other)))

        def __array__(self, dtype=None, copy=None):
            return self._data if dtype is None else self._data.astype(dtype)

        def __len__(self):
            return self._data.shape[0]

        def __repr__(self):
            return f"Statevector({self._data!r}, dims={self._dims})"

    class DensityMatrix:  # noqa: D401 - stand-in
        """Stand-in for ``qiskit.quantum_info.DensityMatrix`` (only what the hot path uses)."""

        def __init__(self, data, dims=None):
            if isinstance(data, Statevector):
                v = data.data
                arr = np.outer(v, v.conj())
            else:
                arr = np.asarray(getattr(data, "data", data), dtype=np.complex128)
                if arr.ndim == 1:
                    arr = np.outer(arr, arr.conj())
            if arr.ndim != 2 or arr.shape[0] != arr.shape[1]:
                raise ValueError("Invalid DensityMatrix input: not a square matrix.")
            self._data = arr
            self._dims = tuple(dims) if dims is not None else _qubit_dims(arr.shape[0])

        @property
        def data(self):
            return self._data

        @property
        def dim(self):
            return self._data.shape[0]

        def dims(self):
            return self._dims

        def trace(self):
            return np.trace(self._data)

        def expectation_value(self, oper):
            return np.trace(np.asarray(oper) @ self._data)

        def probabilities(self):
            return np.real(np.diag(self._data)).copy()

        def probabilities_dict(self):
            n = len(self._dims)
            p = self.probabilities()
            return {format(i, f"0{n}b"): float(v) for i, v in enumerate(p) if v != 0}

        def __array__(self, dtype=None, copy=None):
            return self._data if dtype is None else self._data.astype(dtype)

        def __repr__(self):
            return f"DensityMatrix({self._data!r}, dims={self._dims})"


def is_statevector(obj) -> bool:
    return isinstance(obj, Statevector) or type(obj).__name__ == "Statevector"


def is_density_matrix(obj) -> bool:
    return isinstance(obj, DensityMatrix) or type(obj).__name__ == "DensityMatrix"


__all__ = ["DensityMatrix", "Statevector", "HAVE_QISKIT", "is_statevector", "is_density_matrix"]
