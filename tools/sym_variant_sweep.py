#!/usr/bin/env python
"""Tuning aid for the mirror-symmetric grid kernels: times every kernel variant (forced through the environment
variable BQ_B200_SYM_VARIANT, one subprocess each) on a few (cutoff, grid, frames) cases, next to the library's
own choice.  Prints one JSON line per (case, variant): kernel ms measured with CUDA events around the launch.

    python tools/sym_variant_sweep.py
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [(16, 200, 1), (16, 2048, 1), (32, 400, 64), (32, 2048, 1), (64, 1024, 1), (64, 2048, 1), (128, 2048, 1),
         (256, 2048, 1), (512, 1024, 1), (512, 2048, 1), (1024, 4096, 1)]
RESIDENT = {"auto": None, "mix43x256": 22, "cl4x256": 17, "cl3x128": 8, "cl2x192": 12, "cl2x128": 11, "cl1x256": 13, "cl1x128": 9}
RING = {"auto": None, "cl4x256p8": 21, "cl4x256p": 19, "cl3x384p": 20, "cl4x256r": 18, "cl3x256": 10, "cl2x384": 15, "cl2x256": 14, "cl1x256": 16}


def child():
    sys.path.insert(0, ROOT)
    import torch

    import workloads
    from bosonic_qiskit_b200 import get_engine

    eng = get_engine(0)
    dev = torch.device("cuda", 0)
    tag = os.environ.get("SWEEP_TAG", "auto")
    for N, S, F in CASES:
        names = RESIDENT if N <= 64 else RING
        if tag not in names:
            continue
        rho = torch.from_numpy(workloads.random_rho(N, seed=N, rank=min(N, 4))).to(dev)
        rho = rho.unsqueeze(0).repeat(F, 1, 1).contiguous()
        x = torch.linspace(-6, 6, S, dtype=torch.float64, device=dev)
        out = torch.empty((F, S, S), dtype=torch.float64, device=dev)
        for _ in range(2):
            eng.wigner_t(rho, x, out=out)
        torch.cuda.synchronize()
        eng.profile(True)
        eng.profile_read()
        reps = 5 if N <= 256 else 2
        for _ in range(reps):
            eng.wigner_t(rho, x, out=out)
        ms, n = eng.profile_read()
        eng.profile(False)
        print(json.dumps({"cutoff": N, "grid": S, "frames": F, "variant": tag, "kernel_ms": ms / n}), flush=True)


if __name__ == "__main__":
    if os.environ.get("SWEEP_CHILD"):
        child()
    else:
        only = sys.argv[1].split(",") if len(sys.argv) > 1 else None
        for tag, v in {**RESIDENT, **RING}.items():
            if only and tag not in only:
                continue
            env = dict(os.environ, SWEEP_CHILD="1", SWEEP_TAG=tag)
            if v is not None:
                env["BQ_B200_SYM_VARIANT"] = str(v)
            subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, check=False)
