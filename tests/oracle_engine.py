"""A host-side stand-in for ``bosonic_qiskit_b200.Engine`` backed by the CPU oracle.

TEST INFRASTRUCTURE: lets the ``-m "not gpu"`` suite exercise the drop-in wrappers' host logic (index
mapping, quirks, return types, sharding) in a container without a GPU.  It is never imported by the
package; the GPU suite runs the very same cases against the real CUDA engine.
"""

from __future__ import annotations

import math

import numpy as np

from oracle import partial_trace_dm, partial_trace_sv, wigner_clenshaw


class OracleEngine:
    device = None  # marks it as a host engine for bosonic_qiskit_b200.parallel

    def partial_trace_sv(self, psi, traced):
        psi = np.asarray(psi, dtype=np.complex128)
        if psi.ndim == 1:
            return partial_trace_sv(psi, traced)
        return np.stack([partial_trace_sv(p, traced) for p in psi])

    def partial_trace_sv_multi(self, psi, traced_sets):
        return np.stack([partial_trace_sv(psi, t) for t in traced_sets])

    def partial_trace_dm(self, rho, traced):
        rho = np.asarray(rho, dtype=np.complex128)
        if rho.ndim == 2:
            return partial_trace_dm(rho, traced)
        return np.stack([partial_trace_dm(r, traced) for r in rho])

    def photon_number(self, rho):
        rho = np.asarray(rho, dtype=np.complex128)
        single = rho.ndim == 2
        r3 = rho[None] if single else rho
        n = np.arange(r3.shape[1])
        out = np.array([(n * np.diag(r)).sum() / np.trace(r) for r in r3])
        return out[0] if single else out

    def wigner(self, rho, xvec, yvec=None, g=math.sqrt(2), method="clenshaw", out=None):
        from bosonic_qiskit_b200.engine import method_code

        if method_code(method) == 3:
            return self.wigner_fft(rho, xvec, g)
        rho = np.asarray(rho, dtype=np.complex128)
        if rho.ndim == 2:
            return wigner_clenshaw(rho, xvec, yvec, g)
        return np.stack([wigner_clenshaw(r, xvec, yvec, g) for r in rho])

    def wigner_fft(self, rho, xvec, g=math.sqrt(2)):
        from oracle import wigner_fft

        rho = np.asarray(rho, dtype=np.complex128)
        if rho.ndim == 2:
            return wigner_fft(rho, xvec, g)
        res = [wigner_fft(r, xvec, g) for r in rho]
        return np.stack([r[0] for r in res]), res[0][1]

    def state_to_wigner(self, psi, traced, xvec, yvec=None, g=math.sqrt(2), method="clenshaw", return_rho=False, out=None):
        psi = np.asarray(psi, dtype=np.complex128)
        single = psi.ndim == 1
        p2 = psi[None] if single else psi
        rho = np.stack([partial_trace_sv(p, traced) for p in p2])
        W = np.stack([wigner_clenshaw(r, xvec, yvec, g) for r in rho]) if len(p2) else np.empty((0, len(xvec), len(xvec)))
        if single:
            W, rho = W[0], rho[0]
        return (W, rho) if return_rho else W
