#!/usr/bin/env python
"""Where the time of config c4 goes (4 sites x cutoff 16 from one 16-qubit state, 512 x 512 each): the Wigner kernel from
resident rho, the fused psi -> rho -> W launch, and the unfused two-launch path (BQ_B200_FUSED=0 in a second process),
each timed with CUDA events, L2 flushed between iterations:  python tools/c4_split.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import workloads  # noqa: E402
from bosonic_qiskit_b200 import get_engine  # noqa: E402

# python tools/c4_split.py [sites qubits_per_site grid]   (default: config c4 = 4 sites x 4 qubits, 512 x 512)
if len(sys.argv) > 3:
    import numpy as np

    n_sites, nq, S_ = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    w = dict(psi=workloads.random_state(1 << (n_sites * nq), 0)[None, :], xvec=np.linspace(-6, 6, S_),
             traced_sets=[[q for q in range(n_sites * nq) if not (nq * j <= q < nq * (j + 1))] for j in range(n_sites)])
else:
    w = workloads.config4()
eng = get_engine(0)
dev = torch.device("cuda", 0)
psi = torch.from_numpy(w["psi"][0]).to(dev)
x = torch.from_numpy(w["xvec"]).to(dev)
S = x.numel()
sets = w["traced_sets"]
out = torch.zeros((len(sets), S, S), dtype=torch.float64, device=dev)
D = 1 << (psi.numel().bit_length() - 1 - len(sets[0]))
rho = torch.zeros((len(sets), D, D), dtype=torch.complex128, device=dev)
eng.state_to_wigner_multi_t(psi, sets, x, out=out, rho_out=rho)
flush = torch.empty(1 << 30, dtype=torch.uint8, device=dev)


def timed(fn, reps=20):
    """(microseconds between events around the call with the host kept ahead of the GPU, microseconds of the library's own
    events around its kernel launches)"""
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    eng.profile(True)
    eng.profile_read()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for _ in range(8):
        flush.zero_()
    for a, b in ev:
        if not os.environ.get("NOFLUSH"):
            flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ms, k = eng.profile_read()
    eng.profile(False)
    return round(sum(a.elapsed_time(b) for a, b in ev) / reps * 1e3, 2), round(ms / reps * 1e3, 2), k / reps


res = {
    "fused_env": os.environ.get("BQ_B200_FUSED", "1"), "staged": os.environ.get("BQ_B200_STAGED_TRACE", "1"), "slices": os.environ.get("BQ_B200_PRO_SLICES"), "noflush": os.environ.get("NOFLUSH"),
    "wigner_from_rho_us": timed(lambda: eng.wigner_t(rho, x, out=out)),
    "psi_to_wigner_us": timed(lambda: eng.state_to_wigner_multi_t(psi, sets, x, out=out)),
    "kernel": eng.last_wigner_kernel(),
}
print(json.dumps(res))
