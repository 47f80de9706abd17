"""NumPy restatement of ``qutip.wigner`` (qutip 5.2.2, ``qutip/wigner.py``).

ORACLE -- test infrastructure, not product code (see ``oracle/__init__.py``).

Reference call site: ``src/bosonic_qiskit/wigner.py:190``
``qutip.wigner(psi=qutip.Qobj(rho), xvec=xvec, yvec=yvec, g=g, method=method)``.

The three grid methods QuTiP offers behind ``method=`` are restated
independently (they share no code), so their mutual agreement is itself a
check; the closed forms in ``oracle/analytic.py`` pin the convention
(``W[iy, ix]``, ``g = sqrt(2)`` => coherent |alpha> centred at
``(sqrt2 Re alpha, sqrt2 Im alpha)``).

Operation order inside each function follows the published algorithm
(vectorised over the whole grid, one NumPy temporary per elementary
operation), because this module is also the CPU baseline that ``bench.py``
times: its cost profile should look like the library's, not like a tuned
rewrite.
"""

from __future__ import annotations

import math

import numpy as np
from scipy.special import genlaguerre


def flops_per_point(N: int) -> int:
    """Algorithmic fp64 flops per grid point of the Clenshaw method."""
    return (N - 1) * (5 * N + 6) + 30


def _as_dm(rho) -> np.ndarray:
    a = np.asarray(getattr(rho, "data", rho), dtype=np.complex128)
    if a.ndim == 1:  # ket -> ket2dm (qutip.wigner does this for kets)
        a = np.outer(a, a.conj())
    if a.ndim != 2 or a.shape[0] != a.shape[1]:
        raise TypeError("Input state is not a valid operator.")
    return a


def _laguerre_val(L: int, x: np.ndarray, c: np.ndarray):
    """Clenshaw sum of c_n * (-1)^n sqrt(L! n!/(L+n)!) L_n^L(x) (qutip ``_wig_laguerre_val``)."""
    if len(c) == 1:
        y0 = c[0]
        y1 = 0
    elif len(c) == 2:
        y0 = c[0]
        y1 = c[1]
    else:
        k = len(c)
        y0 = c[-2]
        y1 = c[-1]
        for i in range(3, len(c) + 1):
            k -= 1
            y0, y1 = (
                c[-i] - y1 * (float((k - 1) * (L + k - 1)) / ((L + k) * k)) ** 0.5,
                y0 - y1 * ((L + 2 * k - 1) - x) * ((L + k) * k) ** -0.5,
            )
    return y0 - y1 * ((L + 1) - x) * (L + 1) ** -0.5


def wigner_clenshaw(rho, xvec, yvec=None, g: float = math.sqrt(2)) -> np.ndarray:
    """qutip ``_wigner_clenshaw``: outer Horner in A2/sqrt(L+1), inner Clenshaw in n."""
    rho = _as_dm(rho)
    xvec = np.asarray(xvec, dtype=np.float64)
    yvec = xvec if yvec is None else np.asarray(yvec, dtype=np.float64)
    M = rho.shape[0]
    X, Y = np.meshgrid(xvec, yvec)
    A2 = g * (X + 1.0j * Y)
    B = np.abs(A2)
    B *= B
    w0 = (2 * rho[0, -1]) * np.ones_like(A2)
    L = M - 1
    rho = rho * (2 * np.ones((M, M)) - np.diag(np.ones(M)))
    while L > 0:
        L -= 1
        w0 = _laguerre_val(L, B, np.diag(rho, L)) + w0 * A2 * (L + 1) ** -0.5
    return w0.real * np.exp(-B * 0.5) * (g * g * 0.5 / np.pi)


def wigner_iterative(rho, xvec, yvec=None, g: float = math.sqrt(2)) -> np.ndarray:
    """qutip ``_wigner_iterative``: Wlist[n] recursion over the |m><n| basis functions."""
    rho = _as_dm(rho)
    xvec = np.asarray(xvec, dtype=np.float64)
    yvec = xvec if yvec is None else np.asarray(yvec, dtype=np.float64)
    M = rho.shape[0]
    X, Y = np.meshgrid(xvec, yvec)
    A = 0.5 * g * (X + 1.0j * Y)
    Wlist = np.array([np.zeros(np.shape(A), dtype=complex) for _ in range(M)])
    Wlist[0] = np.exp(-2.0 * np.abs(A) ** 2) / np.pi
    W = np.real(rho[0, 0]) * np.real(Wlist[0])
    for n in range(1, M):
        Wlist[n] = (2.0 * A * Wlist[n - 1]) / np.sqrt(n)
        W += 2 * np.real(rho[0, n] * Wlist[n])
    for m in range(1, M):
        temp = Wlist[m].copy()
        Wlist[m] = (2 * np.conj(A) * temp - np.sqrt(m) * Wlist[m - 1]) / np.sqrt(m)
        W += np.real(rho[m, m] * Wlist[m])
        for n in range(m + 1, M):
            temp2 = (2 * A * Wlist[n - 1] - np.sqrt(m) * temp) / np.sqrt(n)
            temp = Wlist[n].copy()
            Wlist[n] = temp2
            W += 2 * np.real(rho[m, n] * Wlist[n])
    return 0.5 * W * g ** 2


def wigner_laguerre(rho, xvec, yvec=None, g: float = math.sqrt(2)) -> np.ndarray:
    """qutip ``_wigner_laguerre``: explicit associated-Laguerre polynomials per (m, n)."""
    rho = _as_dm(rho)
    xvec = np.asarray(xvec, dtype=np.float64)
    yvec = xvec if yvec is None else np.asarray(yvec, dtype=np.float64)
    M = rho.shape[0]
    X, Y = np.meshgrid(xvec, yvec)
    A = 0.5 * g * (X + 1.0j * Y)
    W = np.zeros(np.shape(A))
    B = 4 * np.abs(A) ** 2
    for m in range(M):
        if abs(rho[m, m]) > 0.0:
            W += np.real(rho[m, m] * (-1) ** m * genlaguerre(m, 0)(B))
        for n in range(m + 1, M):
            if abs(rho[m, n]) > 0.0:
                W += 2.0 * np.real(
                    rho[m, n]
                    * (-1) ** m
                    * (2 * A) ** (n - m)
                    * math.sqrt(math.factorial(m) / math.factorial(n))
                    * genlaguerre(m, n - m)(B)
                )
    return 0.5 * W * g ** 2 * np.exp(-B / 2) / np.pi


_METHODS = {
    "clenshaw": wigner_clenshaw,
    "iterative": wigner_iterative,
    "laguerre": wigner_laguerre,
}


def osc_eigen(N: int, pnts) -> np.ndarray:
    """``qutip.wigner._osc_eigen``: the first N oscillator eigenfunctions at ``pnts`` -> (N, len(pnts))."""
    pnts = np.asarray(pnts, dtype=np.float64)
    A = np.zeros((N, len(pnts)))
    A[0, :] = np.exp(-pnts ** 2 / 2.0) / np.pi ** 0.25
    if N == 1:
        return A
    A[1, :] = np.sqrt(2) * pnts * A[0, :]
    for k in range(2, N):
        A[k, :] = np.sqrt(2.0 / k) * pnts * A[k - 1, :] - np.sqrt((k - 1.0) / k) * A[k - 2, :]
    return A


def _wigner_fft_rows(psi_x: np.ndarray, xvec: np.ndarray):
    """``qutip.wigner._wigner_fft``: Fourier transform of psi(x - s) conj(psi(x + s)) along s for a (1, S) position
    wavefunction; returns (w[x, p], p)."""
    from scipy.linalg import toeplitz

    n = 2 * len(psi_x.T)
    r1 = np.concatenate((np.array([[0]]), np.fliplr(psi_x.conj()), np.zeros((1, n // 2 - 1))), axis=1)
    r2 = np.concatenate((np.array([[0]]), psi_x, np.zeros((1, n // 2 - 1))), axis=1)
    # (scipy <= 1.16, which the reference pins, ravels 2-D c / r; newer scipy treats them as batches: ravel here)
    w = toeplitz(np.zeros(n // 2), r1.ravel()) * np.flipud(toeplitz(np.zeros(n // 2), r2.ravel()))
    w = np.concatenate((w[:, n // 2:n], w[:, 0:n // 2]), axis=1)
    w = np.fft.fft(w)
    w = np.real(np.concatenate((w[:, 3 * n // 4:n + 1], w[:, 0:n // 4]), axis=1))
    p = np.arange(-n / 4, n / 4) * np.pi / (n * (xvec[1] - xvec[0]))
    w = w / (p[1] - p[0]) / n
    return w, p


def wigner_fft(rho, xvec, g: float = math.sqrt(2)):
    """``qutip.wigner(..., method="fft")`` = ``_wigner_fourier``: eigendecompose rho, transform every eigenvector's
    position wavefunction, sum with the eigenvalues.  Returns ``(W[ip, ix], yvec)`` -- the p axis is derived from
    the x grid (``yvec`` of the caller is ignored by this method, as in QuTiP)."""
    rho = np.asarray(rho, dtype=np.complex128)
    xvec = np.asarray(xvec, dtype=np.float64)
    xs = xvec * g / np.sqrt(2)

    def one(psi_col):
        A = osc_eigen(len(psi_col), xs)
        xpsi = np.dot(psi_col.T, A)
        W, y = _wigner_fft_rows(xpsi, xs)
        return (0.5 * g ** 2) * np.real(W.T), y * np.sqrt(2) / g

    if rho.ndim == 1 or 1 in rho.shape:
        return one(rho.reshape(-1, 1))
    vals, vecs = np.linalg.eigh(rho)
    W, yvec = 0, None
    for i in range(rho.shape[0]):
        W1, yvec = one(vecs[:, i].reshape(-1, 1))
        W = W + vals[i] * W1
    return W, yvec


def wigner(rho, xvec, yvec=None, g: float = math.sqrt(2), method: str = "clenshaw") -> np.ndarray:
    """``qutip.wigner(psi, xvec, yvec, method, g)`` for the grid methods.

    Returns ``W[iy, ix]`` float64 of shape ``(len(yvec), len(xvec))``.
    ``method="fft"`` returns the tuple ``(W, yvec)`` on its own p axis (``wigner_fft``).
    """
    if method == "fft":
        return wigner_fft(rho, xvec, g)
    try:
        fn = _METHODS[method]
    except KeyError:
        raise TypeError("method must be either 'iterative', 'laguerre', 'fft' or 'clenshaw'.") from None
    return fn(rho, xvec, yvec, g)
