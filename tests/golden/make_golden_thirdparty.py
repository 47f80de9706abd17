"""Regenerates reference-held golden vectors with the real third-party libraries the reference calls -- qutip
(``qutip.wigner``, src/bosonic_qiskit/wigner.py:190) and qiskit (``qiskit.quantum_info.partial_trace``,
src/bosonic_qiskit/util.py:580,596,619,693).  Neither is installable in the build image (no network), so the committed
fixtures in this directory come from make_golden.py; run THIS script on any machine that has the two libraries

    pip install "qutip>=5.0.4" "qiskit>=2.1,<2.2" numpy scipy
    python tests/golden/make_golden_thirdparty.py

and commit ``thirdparty_c1_c4.npz``: tests/test_oracle.py::test_thirdparty_golden and tests/test_gpu_wigner.py then compare
the oracle and the CUDA engine with it (they skip while the file is absent).
"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import workloads  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    import qiskit
    import qiskit.quantum_info as qi
    import qutip

    g = float(np.sqrt(2))
    c1, c2 = workloads.config1(steps=200), workloads.config2(steps=64)
    c3, c4 = workloads.config3(frames=4, steps=100), workloads.config4(steps=64)
    cases = {"c1": (c1["psi"][0], c1["traced"], c1["xvec"]), "c2": (c2["psi"][0], c2["traced"], c2["xvec"]),
             "c3": (c3["psi"][3], c3["traced"], c3["xvec"]), "c4": (c4["psi"][0], c4["traced_sets"][2], c4["xvec"])}
    out = {"versions": np.array([f"qutip {qutip.__version__}", f"qiskit {qiskit.__version__}", f"numpy {np.__version__}"])}
    for name, (psi, traced, x) in cases.items():
        rho = np.asarray(qi.partial_trace(qi.Statevector(psi), list(traced)).data)
        out[f"{name}_psi"], out[f"{name}_traced"], out[f"{name}_xvec"], out[f"{name}_rho"] = psi, np.array(traced), x, rho
        for method in ("clenshaw", "iterative"):
            out[f"{name}_W_{method}"] = np.asarray(qutip.wigner(qutip.Qobj(rho), x, x, g=g, method=method))
    W, y = qutip.wigner(qutip.Qobj(out["c1_rho"]), c1["xvec"], c1["xvec"], g=g, method="fft")
    out["c1_W_fft"], out["c1_y_fft"] = np.asarray(W), np.asarray(y)
    np.savez_compressed(os.path.join(HERE, "thirdparty_c1_c4.npz"), **out)
    print("written thirdparty_c1_c4.npz with", out["versions"])


if __name__ == "__main__":
    main()
