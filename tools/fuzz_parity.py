#!/usr/bin/env python
"""Randomised parity sweep on a B200: random cutoffs, grid sizes (even / odd, mirror-symmetric or not, square or not),
g, batches and frame counts through the host AND device entry points, against the CPU oracle (TEST INFRASTRUCTURE:
this tool is a checker, it is not imported by the package).  Prints one line per case and a summary; exit code 1 on
the first case above 1e-10 relative to max|W| (or 1e-10 relative to max|rho| for the partial trace).

    python tools/fuzz_parity.py [--cases 300] [--seed 0]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=300)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    import torch

    import workloads
    from bosonic_qiskit_b200 import get_engine
    from oracle import c_oracle, partial_trace_sv, wigner_clenshaw, wigner_fft

    eng = get_engine(0)
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(args.seed)
    worst = {"wigner": 0.0, "trace": 0.0, "fft": 0.0}
    for case in range(args.cases):
        kind = rng.choice(["wigner", "wigner", "wigner", "frames", "fft", "sites"])
        N = int(rng.choice([1, 2, 3, 5, 8, 13, 16, 31, 32, 33, 48, 64, 65, 80, 100, 128, 150, 200]))
        S = int(rng.integers(2, 420))
        g = float(rng.choice([np.sqrt(2), 1.0, 2.0, 0.5 + rng.random()]))
        a = float(2 + 5 * rng.random())
        mirror = bool(rng.random() < 0.6)
        x = np.linspace(-a, a if mirror else a * (0.3 + 0.6 * rng.random()), S)
        if kind == "wigner":
            batch = int(rng.choice([1, 1, 2, 5]))
            rhos = np.stack([workloads.random_rho(N, seed=int(rng.integers(1 << 30)), rank=int(rng.integers(1, 5))) for _ in range(batch)])
            y = x if rng.random() < 0.7 else np.linspace(-a * 0.8, a * 0.9, int(rng.integers(1, 300)))
            eng.set_grid_sharing(bool(rng.random() < 0.8))
            if rng.random() < 0.5:
                W = np.array(eng.wigner(rhos, x, y, g=g))
            else:
                W = eng.wigner_t(torch.from_numpy(rhos).to(dev), torch.from_numpy(x).to(dev),
                                 None if y is x else torch.from_numpy(y).to(dev), g=g).cpu().numpy()
            eng.set_grid_sharing(True)
            ref = np.stack([(c_oracle.wigner_clenshaw(r, x, y, g) if N > 48 else wigner_clenshaw(r, x, y, g)) for r in rhos])
            err = float(np.abs(W - ref).max() / max(np.abs(ref).max(), 1e-300))
            worst["wigner"] = max(worst["wigner"], err)
        elif kind == "frames":
            nq_mode = int(rng.integers(1, 6))
            nq_other = int(rng.integers(0, 4))
            n = nq_mode + nq_other
            F = int(rng.choice([1, 2, 7]))
            psi = rng.normal(size=(F, 1 << n)) + 1j * rng.normal(size=(F, 1 << n))
            psi /= np.linalg.norm(psi, axis=1, keepdims=True)
            traced = sorted(rng.choice(n, size=nq_other, replace=False).tolist())
            W, rho = eng.state_to_wigner(psi, traced, x, g=g, return_rho=True)
            rho_ref = np.stack([partial_trace_sv(p, traced) for p in psi])
            ref = np.stack([wigner_clenshaw(r, x, None, g) for r in rho_ref])
            e_r = float(np.abs(rho - rho_ref).max() / np.abs(rho_ref).max())
            err = float(np.abs(np.array(W) - ref).max() / max(np.abs(ref).max(), 1e-300))
            worst["trace"] = max(worst["trace"], e_r)
            worst["wigner"] = max(worst["wigner"], err)
            err = max(err, e_r)
            N = 1 << nq_mode
        elif kind == "sites":
            # few kept, hundreds to thousands of traced indices at random bit positions (the staged trace of the fused
            # prologue up to 8192 traced indices, the separate trace launch beyond), one or several traces of one state
            nq_mode = int(rng.integers(1, 6))
            nq_other = int(rng.integers(8, 15))
            n = nq_mode + nq_other
            S = min(S, 96)
            x = np.linspace(-a, a if mirror else a * (0.3 + 0.6 * rng.random()), S)
            psi = rng.normal(size=(1 << n)) + 1j * rng.normal(size=(1 << n))
            psi /= np.linalg.norm(psi)
            n_sets = int(rng.choice([1, 1, 3]))
            sets = [sorted(rng.choice(n, size=nq_other, replace=False).tolist()) for _ in range(n_sets)]
            if n_sets == 1:
                W, rho = eng.state_to_wigner(psi[None, :], sets[0], x, g=g, return_rho=True)
            else:
                W, rho = eng.state_to_wigner_multi(psi, sets, x, g=g, return_rho=True)
            rho_ref = np.stack([partial_trace_sv(psi, t) for t in sets])
            ref = np.stack([wigner_clenshaw(r, x, None, g) for r in rho_ref])
            e_r = float(np.abs(np.array(rho).reshape(rho_ref.shape) - rho_ref).max() / np.abs(rho_ref).max())
            err = float(np.abs(np.array(W).reshape(ref.shape) - ref).max() / max(np.abs(ref).max(), 1e-300))
            worst["trace"] = max(worst["trace"], e_r)
            worst["wigner"] = max(worst["wigner"], err)
            err = max(err, e_r)
            N = 1 << nq_mode
        else:
            N = min(N, 64)
            S = min(S, 200)
            x = np.linspace(-a, a * (1 if mirror else 0.7), S)
            rho = workloads.random_rho(N, seed=int(rng.integers(1 << 30)), rank=2)
            W, yv = eng.wigner_fft(rho, x, g=g)
            ref, yr = wigner_fft(rho, x, g)
            err = float(np.abs(W - ref).max() / max(np.abs(ref).max(), 1e-300))
            err = max(err, float(np.abs(yv - yr).max() / np.abs(yr).max()) * 1e-3)
            worst["fft"] = max(worst["fft"], err)
        print(f"{case:4d} {kind:7s} N={N:4d} S={S:4d} g={g:.3f} mirror={int(mirror)} err={err:.2e}", flush=True)
        if not err < 1e-10:
            print("FAIL")
            sys.exit(1)
    print("FUZZ PASS", args.cases, "cases; worst relative errors:", {k: f"{v:.2e}" for k, v in worst.items()})


if __name__ == "__main__":
    main()
