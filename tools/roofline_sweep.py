#!/usr/bin/env python
"""fp64 roofline sweep of the Wigner kernel: random rho at several cutoffs, device-resident, CUDA events.

    python tools/roofline_sweep.py [--grid 2048] [--cutoffs 16,32,64,128,256,512,1024]
Prints one JSON line per cutoff: points/s, achieved TFLOP/s (algorithmic flops of SURVEY 8d) and the
fraction of the fp64 peak measured in the same process.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import workloads  # noqa: E402
from bosonic_qiskit_b200 import get_engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=2048)
    ap.add_argument("--cutoffs", default="16,32,64,128,256,512,1024")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--sass", default="patched", choices=["patched", "plain"])
    ap.add_argument("--sharing", default="off", choices=["on", "off"],
                    help="on: mirror-symmetric grid kernels (one recurrence per pair of grid lines); "
                         "off: point-by-point kernels (the fp64 roofline measurement)")
    args = ap.parse_args()
    eng = get_engine(0)
    dev = torch.device("cuda", 0)
    active = eng.set_sass_mode(args.sass == "patched")
    eng.set_grid_sharing(args.sharing == "on")
    peak = eng.fp64_peak_tflops(1.0)
    print(json.dumps({"fp64_peak_tflops_measured": peak, "sass": "patched" if active else "plain",
                      "sharing": args.sharing}), flush=True)
    for N in [int(c) for c in args.cutoffs.split(",")]:
        S = args.grid if N <= 256 else max(512, args.grid // (2 if N == 512 else 4))
        rho = torch.from_numpy(workloads.random_rho(N, seed=N, rank=min(N, 8))).to(dev)
        x = torch.linspace(-6, 6, S, dtype=torch.float64, device=dev)
        out = torch.empty((1, S, S), dtype=torch.float64, device=dev)
        for _ in range(2):
            eng.wigner_t(rho, x, out=out)
        torch.cuda.synchronize()
        eng.profile(True)
        eng.profile_read()
        reps = args.reps if N <= 256 else 2
        for _ in range(reps):
            eng.wigner_t(rho, x, out=out)
        ms, n = eng.profile_read()
        eng.profile(False)
        ms /= n
        pts = S * S
        tf = pts * workloads.flops_per_point(N) / (ms * 1e-3) / 1e12
        print(json.dumps({"cutoff": N, "grid": [S, S], "kernel_ms": ms, "points_per_s": pts / (ms * 1e-3),
                          "algorithmic_tflops": tf, "algorithmic_frac_of_measured_fp64_peak": tf / peak}), flush=True)


if __name__ == "__main__":
    main()
