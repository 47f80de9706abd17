#!/usr/bin/env python
"""Kernel timing of the Wigner launch on a few (cutoff, grid, frames) cases; CUDA events around the launch
(bq_profile_enable).  Environment switches of the library (BQ_B200_*) apply; one JSON line per case.

    python tools/ktime.py [N:S:F ...]      default: 64:1024:1 1024:4096:1
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import workloads  # noqa: E402
from bosonic_qiskit_b200 import get_engine  # noqa: E402


def main():
    cases = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:] if ":" in a] or [(64, 1024, 1), (1024, 4096, 1)]
    eng = get_engine(0)
    dev = torch.device("cuda", 0)
    tag = os.environ.get("KTIME_TAG", "")
    peak = eng.fp64_peak_tflops(0.5)
    for N, S, F in cases:
        rho = torch.from_numpy(workloads.random_rho(N, seed=N, rank=min(N, 4))).to(dev)
        rho = rho.unsqueeze(0).repeat(F, 1, 1).contiguous()
        x = torch.linspace(-6, 6, S, dtype=torch.float64, device=dev)
        out = torch.empty((F, S, S), dtype=torch.float64, device=dev)
        for _ in range(3):
            eng.wigner_t(rho, x, out=out)
        torch.cuda.synchronize()
        eng.profile(True)
        eng.profile_read()
        reps = 10 if N <= 256 else 3
        for _ in range(reps):
            eng.wigner_t(rho, x, out=out)
        ms, n = eng.profile_read()
        eng.profile(False)
        ms /= reps
        units = F * workloads.shared_units(S)
        ex = units * workloads.flops_per_unit_shared(N) / (ms * 1e-3) / 1e12
        print(json.dumps({"tag": tag, "cutoff": N, "grid": S, "frames": F, "kernel": eng.last_wigner_kernel(), "kernel_ms": round(ms, 5),
                          "launches_per_call": n / reps, "executed_frac": round(ex / peak, 4), "checksum": float(out.sum().item())}), flush=True)


if __name__ == "__main__":
    main()
