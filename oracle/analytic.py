"""Closed-form Wigner functions and deterministic synthetic states.

ORACLE -- test infrastructure, not product code (see ``oracle/__init__.py``).

Convention pinned by the reference's textbook notebook
(``tutorials/bosonic-qiskit-textbook/4. wigner_distribution.ipynb`` cells
17-19): with ``g = sqrt(2)`` a coherent state |alpha> is a Gaussian centred at
``(x, p) = (sqrt2 Re alpha, sqrt2 Im alpha)`` with variance 1/2 per axis, and
the returned array is indexed ``W[iy, ix]`` (numpy ``meshgrid`` 'xy').

For general ``g`` substitute ``beta = g (x + i p) / 2``:
``W_g(x, p) = (g^2/4) * (2/pi) * <D(beta) Pi D(beta)^dagger>`` evaluated below
through the g = sqrt(2) closed forms with ``(x, p) -> g (x, p) / sqrt2`` and an
overall factor ``g^2 / 2``.
"""

from __future__ import annotations

import math

import numpy as np
from scipy.special import eval_genlaguerre, gammaln


def _grid(xvec, yvec, g):
    xvec = np.asarray(xvec, dtype=np.float64)
    yvec = xvec if yvec is None else np.asarray(yvec, dtype=np.float64)
    X, Y = np.meshgrid(xvec, yvec)
    s = g / math.sqrt(2.0)
    return X * s, Y * s, 0.5 * g * g


def wigner_fock(n: int, xvec, yvec=None, g: float = math.sqrt(2)) -> np.ndarray:
    """|n><n|: W = (1/pi) (-1)^n L_n(2 r^2) exp(-r^2)."""
    X, Y, sc = _grid(xvec, yvec, g)
    r2 = X * X + Y * Y
    return sc * ((-1) ** n / np.pi) * eval_genlaguerre(n, 0, 2 * r2) * np.exp(-r2)


def wigner_coherent(alpha: complex, xvec, yvec=None, g: float = math.sqrt(2)) -> np.ndarray:
    """|alpha><alpha| (untruncated): Gaussian at sqrt2 (Re alpha, Im alpha)."""
    X, Y, sc = _grid(xvec, yvec, g)
    x0, p0 = math.sqrt(2) * alpha.real, math.sqrt(2) * alpha.imag
    return sc * np.exp(-((X - x0) ** 2) - (Y - p0) ** 2) / np.pi


def wigner_element(m: int, n: int, xvec, yvec=None, g: float = math.sqrt(2)) -> np.ndarray:
    """Contribution of a Hermitian pair rho = |m><n| + |n><m| (m < n), or |m><m| for m == n.

    W = (2/pi)(-1)^m sqrt(m!/n!) Re[(2 beta)^(n-m)] ... with beta = (x+ip)/sqrt2; this is
    the term the 'laguerre' method sums, written with log-gamma for large indices.
    """
    if m > n:
        m, n = n, m
    X, Y, sc = _grid(xvec, yvec, g)
    A = (X + 1j * Y) / math.sqrt(2)  # = g(x+ip)/2 at g = sqrt2
    B = 4 * np.abs(A) ** 2
    lag = eval_genlaguerre(m, n - m, B)
    pref = math.exp(0.5 * (gammaln(m + 1) - gammaln(n + 1)))
    core = ((-1) ** m) * pref * ((2 * A) ** (n - m)) * lag
    w = np.real(core) * (1.0 if m == n else 2.0)
    return sc * w * np.exp(-B / 2) / np.pi


def wigner_cat(alpha: float, parity: int, xvec, yvec=None, g: float = math.sqrt(2)) -> np.ndarray:
    """Even (+1) / odd (-1) cat (|alpha> +- |-alpha>)/norm for REAL alpha, untruncated."""
    X, Y, sc = _grid(xvec, yvec, g)
    a = math.sqrt(2) * alpha
    norm = 2.0 * (1.0 + parity * math.exp(-2.0 * alpha * alpha))
    gp = np.exp(-((X - a) ** 2) - Y ** 2)
    gm = np.exp(-((X + a) ** 2) - Y ** 2)
    inter = 2.0 * parity * np.exp(-(X ** 2) - Y ** 2) * np.cos(2.0 * a * Y)
    return sc * (gp + gm + inter) / (np.pi * norm)


# ---------------------------------------------------------------------------
# deterministic synthetic states; pure NumPy/SciPy, no qiskit
# ---------------------------------------------------------------------------

def coherent_amplitudes(alpha: complex, N: int) -> np.ndarray:
    """Truncated, renormalised coherent-state amplitudes c_n = e^{-|a|^2/2} a^n / sqrt(n!)."""
    n = np.arange(N)
    logc = -0.5 * abs(alpha) ** 2 - 0.5 * gammaln(n + 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        c = np.exp(logc + n * np.log(abs(alpha) if alpha != 0 else 1.0)) * np.exp(1j * n * np.angle(alpha))
    if alpha == 0:
        c = np.zeros(N, dtype=complex)
        c[0] = 1.0
    return (c / np.linalg.norm(c)).astype(np.complex128)


def cat_amplitudes(alpha: complex, N: int, parity: int = +1) -> np.ndarray:
    c = coherent_amplitudes(alpha, N) + parity * coherent_amplitudes(-alpha, N)
    return c / np.linalg.norm(c)


def displacement_matrix(alpha: complex, N: int) -> np.ndarray:
    """expm(alpha a^dagger - conj(alpha) a) on the truncated space.

    Same construction as the reference's ``CVOperators.d``
    (``src/bosonic_qiskit/operators.py:142-153``): truncated ladder operators,
    then a matrix exponential of the generator.
    """
    from scipy.linalg import expm

    a = np.diag(np.sqrt(np.arange(1, N)), 1).astype(np.complex128)
    return expm(alpha * a.conj().T - np.conj(alpha) * a)


def random_state(dim: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(dim) + 1j * rng.standard_normal(dim)
    return (v / np.linalg.norm(v)).astype(np.complex128)


def random_density_matrix(N: int, seed: int, rank: int | None = None) -> np.ndarray:
    rng = np.random.default_rng(seed)
    r = N if rank is None else rank
    G = rng.standard_normal((N, r)) + 1j * rng.standard_normal((N, r))
    rho = G @ G.conj().T
    return (rho / np.trace(rho).real).astype(np.complex128)


def jc_frames(F: int, N: int, alpha: float = 2.0, omega_r: float = 2.0, chi: float = 0.1) -> np.ndarray:
    """c3: F frames of a dispersive Jaynes-Cummings evolution, one qubit in |0>.

    psi_k[q*N + n] (q = qubit bit above the qumode bits): coherent |alpha> whose
    Fock components rotate at (omega_r + chi/2).  Returns (F, 2N) complex128.
    """
    n = np.arange(N)
    c = coherent_amplitudes(alpha, N)
    t = np.arange(F) * (2 * np.pi / omega_r) / F
    ph = np.exp(-1j * np.outer(t, n) * (omega_r + chi / 2))
    out = np.zeros((F, 2 * N), dtype=np.complex128)
    out[:, :N] = c[None, :] * ph
    return out
