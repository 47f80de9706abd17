"""NumPy restatement of what ``animate._animate`` / ``wigner.plot`` make matplotlib do with a Wigner frame
(``src/bosonic_qiskit/animate.py:278-291``, ``wigner.py:260-269``): symmetric colour levels
``linspace(-abs_max, abs_max, 100)`` and ``contourf(..., levels, cmap="RdBu")``, evaluated at the grid nodes.

ORACLE -- test infrastructure, not product code.  matplotlib is not installed in this image, so this restates its
published behaviour (3.x): ``ContourSet`` colours layer ``j`` (between ``levels[j]`` and ``levels[j+1]``) with
``cmap(norm(0.5 * (levels[j] + levels[j+1])))``, ``norm = Normalize(levels[0], levels[-1])``; ``RdBu`` is
``LinearSegmentedColormap.from_list`` over ColorBrewer's 11 anchors with a 256-entry lookup table;
``Colormap.__call__(bytes=True)`` indexes it with ``int(x * 256)`` (256 -> 255) and returns ``(lut * 255).astype(uint8)``.
PARITY UNPINNED against matplotlib itself (absent here); pinned only against these published rules.
"""

from __future__ import annotations

import numpy as np

# ColorBrewer RdBu, 11 classes (matplotlib/_cm.py::_RdBu_data), as n / 255
_RDBU = np.array([(103, 0, 31), (178, 24, 43), (214, 96, 77), (244, 165, 130), (253, 219, 199), (247, 247, 247),
                  (209, 229, 240), (146, 197, 222), (67, 147, 195), (33, 102, 172), (5, 48, 97)], dtype=np.float64) / 255.0


def rdbu_lut(n: int = 256) -> np.ndarray:
    """``LinearSegmentedColormap.from_list("RdBu", _RdBu_data, N=256)._lut[:256]`` as uint8 RGBA (alpha 255)."""
    x = np.linspace(0, 1, len(_RDBU))
    xind = np.linspace(0, 1, n)
    lut = np.empty((n, 4))
    for ch in range(3):
        y = _RDBU[:, ch]
        ind = np.searchsorted(x, xind)[1:-1]
        distance = (xind[1:-1] - x[ind - 1]) / (x[ind] - x[ind - 1])
        lut[:, ch] = np.clip(np.concatenate([[y[0]], distance * (y[ind] - y[ind - 1]) + y[ind - 1], [y[-1]]]), 0, 1)
    lut[:, 3] = 1.0
    return (lut * 255).astype(np.uint8)


def frame_rgba(W: np.ndarray, levels: int = 100, zero_abs_max: float = 5.0) -> np.ndarray:
    """(S, S) float64 -> (S, S, 4) uint8: the colour ``contourf(xvec, xvec, W, linspace(-A, A, levels), cmap="RdBu")``
    gives each grid node, ``A = max(amax, |amin|)`` (``zero_abs_max`` if that is 0: 5 in ``_animate``, 1 in ``plot``)."""
    W = np.asarray(W, dtype=np.float64)
    amax, amin = np.max(W), np.min(W)
    A = max(amax, abs(amin))
    if A == 0:
        A = zero_abs_max
    lev = np.linspace(-A, A, levels)
    band = np.clip(np.searchsorted(lev, W, side="right") - 1, 0, levels - 2)
    layer = 0.5 * (lev[:-1] + lev[1:])
    norm = (layer - lev[0]) / (lev[-1] - lev[0])
    idx = (norm * 256).astype(int)
    idx[idx == 256] = 255
    return rdbu_lut()[idx[band]]
