"""Deterministic synthetic inputs for BASELINE.json's five configs.

Input generation only: pure NumPy/SciPy, no qiskit, no oracle, no GPU.  The states are built with
the same constructions the reference's gate library uses (truncated ladder operators and a matrix
exponential of the generator: ``src/bosonic_qiskit/operators.py:142-153`` displacement,
``:207-219`` beam splitter, ``:257-270`` conditional displacement), laid out in the reference's
index convention (``qumoderegister.py:85-88``, ``circuit.py:50-72``): registers in constructor
order, qumode k on consecutive qubits, Fock number little-endian.

Every workload is a dict:
    psi      (frames, 2**n_qubits) complex128   statevectors (frames >= 1)
    n_qubits int
    traced   list[int]  or  traced_sets list[list[int]] (several traces of one psi: config 4)
    N        int        cutoff of the reduced state (2**kept qubits)
    xvec     (S,) float64, grid is xvec x xvec
    name     str
"""

from __future__ import annotations

import numpy as np
from scipy.sparse import diags, identity, kron
from scipy.sparse.linalg import expm_multiply
from scipy.special import gammaln

G_DEFAULT = float(np.sqrt(2))


def _ladder(N: int):
    return diags(np.sqrt(np.arange(1, N)), 1, format="csc", dtype=np.complex128)


def coherent(alpha: complex, N: int) -> np.ndarray:
    """Truncated + renormalised coherent amplitudes e^{-|a|^2/2} a^n / sqrt(n!)."""
    n = np.arange(N)
    if alpha == 0:
        c = np.zeros(N, dtype=np.complex128)
        c[0] = 1
        return c
    logmag = -0.5 * abs(alpha) ** 2 + n * np.log(abs(alpha)) - 0.5 * gammaln(n + 1)
    c = np.exp(logmag) * np.exp(1j * n * np.angle(alpha))
    return (c / np.linalg.norm(c)).astype(np.complex128)


def displace(alpha: complex, vec: np.ndarray) -> np.ndarray:
    """expm(alpha a^dag - conj(alpha) a) @ vec on the truncated space (operators.py:142-153)."""
    N = vec.shape[0]
    a = _ladder(N)
    gen = (alpha * a.conj().T - np.conj(alpha) * a).tocsc()
    return expm_multiply(gen, vec.astype(np.complex128))


def random_state(dim: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(dim) + 1j * rng.standard_normal(dim)
    return (v / np.linalg.norm(v)).astype(np.complex128)


def random_rho(N: int, seed: int, rank: int | None = None) -> np.ndarray:
    rng = np.random.default_rng(seed)
    r = N if rank is None else rank
    G = rng.standard_normal((N, r)) + 1j * rng.standard_normal((N, r))
    rho = G @ G.conj().T
    return (rho / np.trace(rho).real).astype(np.complex128)


# ------------------------------------------------------------------------------------------------
def config1(steps: int = 200) -> dict:
    """1 qumode cutoff 16 (4 qubits) + 1 qubit, displaced cat alpha=2, 200x200 grid."""
    N, alpha = 16, 2.0
    vac = np.zeros(N, dtype=np.complex128)
    vac[0] = 1
    dp, dm = displace(alpha, vac), displace(-alpha, vac)
    psi = np.zeros(2 * N, dtype=np.complex128)
    psi[:N] = 0.5 * (dp + dm)  # qubit |0>  (index = q*16 + n)
    psi[N:] = 0.5 * (dp - dm)  # qubit |1>
    return dict(name="c1", psi=psi[None, :], n_qubits=5, traced=[4], N=N, xvec=np.linspace(-6, 6, steps))


def config2(steps: int = 1024) -> dict:
    """1 qubit + 2 qumodes cutoff 64: conditional displacement (alpha=1.5 on mode 0) then a
    beamsplitter (theta=pi/4); trace to qumode 0 -> 64x64 rho; 1024x1024 grid.
    Index = q*4096 + n1*64 + n0 (registers: qumodes first, then the qubit)."""
    N, alpha, theta = 64, 1.5, np.pi / 4
    vac = np.zeros(N, dtype=np.complex128)
    vac[0] = 1
    a = _ladder(N)
    eye = identity(N, format="csc", dtype=np.complex128)
    a0 = kron(eye, a, format="csc")  # mode 0 = least-significant factor
    a1 = kron(a, eye, format="csc")
    gen = (theta * (a0.conj().T @ a1) - np.conj(theta) * (a1.conj().T @ a0)).tocsc()
    psi = np.zeros(2 * N * N, dtype=np.complex128)
    for q, sgn in ((0, +1), (1, -1)):
        m0 = displace(sgn * alpha, vac)  # cD: D(alpha) for |0>, D(-alpha) for |1>
        two = np.kron(vac, m0)  # |n1=0> (x) |m0>, index n1*64 + n0
        psi[q * N * N:(q + 1) * N * N] = expm_multiply(gen, two) / np.sqrt(2)  # qubit in |+>
    traced = list(range(6, 13))  # mode 1 (bits 6..11) and the qubit (bit 12)
    return dict(name="c2", psi=psi[None, :], n_qubits=13, traced=traced, N=N, xvec=np.linspace(-6, 6, steps))


def config3(frames: int = 512, steps: int = 400) -> dict:
    """Jaynes-Cummings (dispersive) animation: `frames` frames, cutoff 32 + 1 qubit in |0>."""
    N, alpha, omega_r, chi = 32, 2.0, 2.0, 0.1
    n = np.arange(N)
    c = coherent(alpha, N)
    t = np.arange(frames) * (2 * np.pi / omega_r) / 512
    ph = np.exp(-1j * np.outer(t, n) * (omega_r + chi / 2))
    psi = np.zeros((frames, 2 * N), dtype=np.complex128)
    psi[:, :N] = c[None, :] * ph
    return dict(name="c3", psi=psi, n_qubits=6, traced=[5], N=N, xvec=np.linspace(-6, 6, steps))


def config4(steps: int = 512, seed: int = 0) -> dict:
    """Bose-Hubbard-like register: 4 sites x cutoff 16 (16 qubits), random normalised state;
    per-site reduced density matrices (4 traces of one psi) + 512x512 Wigner each."""
    n_sites, nq = 4, 4
    psi = random_state(1 << (n_sites * nq), seed)
    sets = [[q for q in range(n_sites * nq) if not (nq * j <= q < nq * (j + 1))] for j in range(n_sites)]
    return dict(name="c4", psi=psi[None, :], n_qubits=n_sites * nq, traced_sets=sets, N=16,
                xvec=np.linspace(-6, 6, steps))


def config5(steps: int = 4096, kind: str = "cat", N: int = 1024) -> dict:
    """Single qumode cutoff 1024 (10 qubits), no qubit to trace; 4096x4096 grid (row-slab sharded)."""
    if kind == "cat":
        c = coherent(4.0, N) + coherent(-4.0, N)
        psi = c / np.linalg.norm(c)
    else:
        psi = random_state(N, 1)
    return dict(name="c5", psi=psi[None, :], n_qubits=int(np.log2(N)), traced=[], N=N, xvec=np.linspace(-6, 6, steps))


CONFIGS = {"c1": config1, "c2": config2, "c3": config3, "c4": config4, "c5": config5}


def flops_per_point(N: int) -> int:
    """Algorithmic fp64 flops per Wigner grid point at cutoff N."""
    return (N - 1) * (5 * N + 6) + 30


def shared_units(S: int) -> int:
    """Pairs of mirror classes of an S x S mirror-symmetric grid: one Clenshaw recurrence each (wigner_sym_* kernels)."""
    U = (S + 1) // 2
    return U * (U + 1) // 2


def flops_per_unit_shared(N: int) -> int:
    """fp64 flops the symmetric-grid kernels execute per unit (FMA = 2): 10 per inner step, and per diagonal the
    closing S_L (6), the scaled A2 (2) and eight complex Horner updates (8 x 8); ~40 for exp / scale / B."""
    return (N - 1) * (5 * (N - 2) + 72) + 40

