#!/usr/bin/env python
"""Tuning aid: times the mirror-symmetric grid kernels as ONE launch (the library's choice) against TWO launches
(whole rounds with variant `main`, the remainder with variant `tail`; forced through BQ_B200_SYM_SPLIT=main:tail,
one subprocess each).  One JSON line per (case, split): kernel ms with CUDA events around the launch pair.

    python tools/sym_split_sweep.py
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [(16, 200, 1), (16, 1024, 1), (32, 400, 8), (32, 1024, 1), (64, 512, 1), (64, 700, 1), (64, 1024, 1), (64, 1400, 1),
         (64, 2048, 1), (128, 1024, 1), (256, 1024, 1), (256, 2048, 1)]
SPLITS = ["", "8:9", "8:13", "8:11", "11:9", "11:13", "10:14", "10:16", "14:16"]


def child():
    sys.path.insert(0, ROOT)
    import torch

    import workloads
    from bosonic_qiskit_b200 import get_engine

    eng = get_engine(0)
    dev = torch.device("cuda", 0)
    tag = os.environ.get("BQ_B200_SYM_SPLIT", "")
    for N, S, F in CASES:
        resident = N <= 64
        if tag and (int(tag.split(":")[0]) in (10, 14, 15)) == resident:
            continue
        rho = torch.from_numpy(workloads.random_rho(N, seed=N, rank=min(N, 4))).to(dev)
        rho = rho.unsqueeze(0).repeat(F, 1, 1).contiguous()
        x = torch.linspace(-6, 6, S, dtype=torch.float64, device=dev)
        out = torch.empty((F, S, S), dtype=torch.float64, device=dev)
        try:
            for _ in range(2):
                eng.wigner_t(rho, x, out=out)
        except Exception as exc:
            print(json.dumps({"cutoff": N, "grid": S, "split": tag, "error": str(exc)}), flush=True)
            continue
        torch.cuda.synchronize()
        l0 = eng.launches
        eng.profile(True)
        eng.profile_read()
        reps = 5
        for _ in range(reps):
            eng.wigner_t(rho, x, out=out)
        ms, n = eng.profile_read()
        eng.profile(False)
        print(json.dumps({"cutoff": N, "grid": S, "frames": F, "split": tag or "one launch", "kernel_ms": round(ms / n, 5),
                          "launches_per_call": (eng.launches - l0) / reps}), flush=True)


if __name__ == "__main__":
    if os.environ.get("SWEEP_CHILD"):
        child()
    else:
        for tag in SPLITS:
            env = dict(os.environ, SWEEP_CHILD="1")
            env.pop("BQ_B200_SYM_SPLIT", None)
            if tag:
                env["BQ_B200_SYM_SPLIT"] = tag
            subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, check=False)
