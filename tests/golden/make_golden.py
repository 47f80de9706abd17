"""Regenerates the committed golden fixtures in this directory.

The reference cannot be imported in the build image (qutip / qiskit / qiskit-aer are absent), so the
fixtures are (i) the values PRINTED in the reference's own notebook
(`tutorials/wigner-function/wigner-function.ipynb:309-313,353-361`, transcribed below) and
(ii) outputs of the CPU oracle for BASELINE.json's config 1, after that oracle has been pinned to
the analytic closed forms by tests/test_oracle.py.  Run from the repo root:

    python tests/golden/make_golden.py
"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import analytic as an  # noqa: E402
from oracle import partial_trace_sv, wigner_clenshaw  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def notebook():
    # cutoff 4 (2 qubits) + 1 qubit in |0>, after cv_d(i*pi/2) on the vacuum
    psi = np.array([0.27875462, 0.51187782j, -0.36358351, -0.72669388j, 0, 0, 0, 0], dtype=np.complex128)
    rho = np.array(
        [
            [0.07770414, -0.14268831j, -0.10135058, 0.20256928j],
            [0.14268831j, 0.2620189, -0.18611034j, -0.37197848],
            [-0.10135058, 0.18611034j, 0.13219297, -0.26421391j],
            [-0.20256928j, -0.37197848, 0.26421391j, 0.52808399],
        ],
        dtype=np.complex128,
    )
    np.savez(os.path.join(HERE, "notebook_cutoff4.npz"), psi=psi, rho=rho, traced=np.array([2]))


def config1():
    # BASELINE.json configs[0]: 1 qumode cutoff 16 (4 qubits) + 1 qubit, displaced cat alpha=2
    N, alpha = 16, 2.0
    Dp, Dm = an.displacement_matrix(alpha, N), an.displacement_matrix(-alpha, N)
    vac = np.zeros(N, dtype=np.complex128)
    vac[0] = 1.0
    plus, minus = (Dp + Dm) @ vac, (Dp - Dm) @ vac
    psi = np.zeros(2 * N, dtype=np.complex128)
    psi[:N] = 0.5 * plus  # qubit |0>
    psi[N:] = 0.5 * minus  # qubit |1>
    rho = partial_trace_sv(psi, [4])
    xvec = np.linspace(-6, 6, 200)
    W = wigner_clenshaw(rho, xvec)
    np.savez_compressed(os.path.join(HERE, "config1_cat.npz"), psi=psi, rho=rho, W=W, traced=np.array([4]))


if __name__ == "__main__":
    notebook()
    config1()
    print("written:", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
