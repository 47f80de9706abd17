import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import workloads
from bosonic_qiskit_b200 import get_engine
eng = get_engine(0)
wl = workloads.config2()
S = len(wl["xvec"])
psi_pin = eng.pinned_empty(wl["psi"].shape, np.complex128); psi_pin[...] = wl["psi"]
out_pin = eng.pinned_empty((1, S, S), np.float64)
out_page = np.empty((1, S, S))
for name, psi, out in (("pinned in/out (zero-copy store)", psi_pin, out_pin), ("pageable out (staged D2H)", psi_pin, out_page)):
    for _ in range(5): eng.state_to_wigner(psi, wl["traced"], wl["xvec"], out=out)
    t0 = time.perf_counter(); dev = 0.0
    for _ in range(50):
        eng.state_to_wigner(psi, wl["traced"], wl["xvec"], out=out); dev += eng.last_device_ms()
    wall = (time.perf_counter() - t0) / 50 * 1e3
    print(f"{name}: wall {wall:.3f} ms/call, device span {dev/50:.3f} ms")
# device-resident for reference
import torch
psi_t = torch.from_numpy(wl["psi"]).cuda(); x_t = torch.from_numpy(wl["xvec"]).cuda(); W_t = torch.empty((1,S,S), dtype=torch.float64, device='cuda')
for _ in range(5): eng.state_to_wigner_t(psi_t, wl["traced"], x_t, out=W_t)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50): eng.state_to_wigner_t(psi_t, wl["traced"], x_t, out=W_t)
torch.cuda.synchronize(); print(f"device-resident back-to-back: {(time.perf_counter()-t0)/50*1e3:.3f} ms/call")
# zero-copy output from the dev API: does storing to host slow the kernel?
import ctypes
eng.profile(True); eng.profile_read()
for _ in range(20): eng.state_to_wigner(psi_pin, wl["traced"], wl["xvec"], out=out_pin)
ms, n = eng.profile_read(); print(f"wigner kernel with host-mapped output: {ms/n:.4f} ms")
for _ in range(20): eng.state_to_wigner_t(psi_t, wl["traced"], x_t, out=W_t)
torch.cuda.synchronize(); ms, n = eng.profile_read(); print(f"wigner kernel with device output:      {ms/n:.4f} ms")
