#!/usr/bin/env python
"""Kernel time of ONE GPU's share of a large mirror-symmetric grid (unit range = 1/parts of the grid's units), i.e. what each
rank of the strong-scaling bench runs: python tools/share_time.py [N S parts ...]   (default 1024 4096 1,2,4,8)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import workloads  # noqa: E402
from bosonic_qiskit_b200 import get_engine  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
parts = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "1,2,4,8").split(",")]
eng = get_engine(0)
dev = torch.device("cuda", 0)
rho = torch.from_numpy(workloads.random_rho(N, seed=N, rank=4)).to(dev).unsqueeze(0)
x = torch.linspace(-6, 6, S, dtype=torch.float64, device=dev)
out = torch.zeros((1, S, S), dtype=torch.float64, device=dev)
units = eng.grid_units(S)
peak = eng.fp64_peak_tflops(0.3)
for n in parts:
    cnt = units // n
    for _ in range(2):
        eng.wigner_t(rho, x, out=out, unit_range=(0, cnt))
    torch.cuda.synchronize()
    eng.profile(True)
    eng.profile_read()
    for _ in range(3):
        eng.wigner_t(rho, x, out=out, unit_range=(0, cnt))
    ms, k = eng.profile_read()
    eng.profile(False)
    ms /= 3
    ex = cnt * workloads.flops_per_unit_shared(N) / (ms * 1e-3) / 1e12
    print(json.dumps({"cutoff": N, "grid": S, "parts": n, "units": cnt, "kernel_ms": round(ms, 3), "launches": k / 3,
                      "last_kernel": eng.last_wigner_kernel(), "executed_frac": round(ex / peak, 4), "times_parts_ms": round(ms * n, 2)}), flush=True)
