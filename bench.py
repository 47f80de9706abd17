#!/usr/bin/env python
"""Benchmark of the statevector -> reduced rho -> Wigner grid hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c2]

One "step" is one pass of the hot path over one batch of synthetic input: for the default config
(BASELINE.json configs[1], "c2": 1 qubit + 2 qumodes at cutoff 64, trace to qumode 0, 1024 x 1024
Wigner grid) that is  psi (8192 amplitudes) -> rho (64 x 64) -> W (1024 x 1024 float64).

Under torchrun (N > 1) every rank runs the same per-rank workload (weak scaling: the frames of a
batch are independent and are sharded across ranks), then the finished grids are gathered on rank 0
over NCCL inside the timed step.  Rank 0 prints ONE JSON line.

  value     whole-job Wigner grid points / s, inputs resident in HBM (CUDA events, max over ranks)
  e2e       same metric through the public NumPy API (C-ABI `_host` entry point): pinned host psi
            -> H2D -> kernels -> D2H -> host W inside the timed region (wall clock)
  roofline  dominant kernel (the Wigner Clenshaw kernel): algorithmic fp64 flops / its own event
            time, against an fp64-FMA peak MEASURED on this GPU in the same run (MEASURED_PEAKS.json
            has no fp64 entry); the HBM store side is reported next to it against hbm_gbs
  cpu_baseline  the NumPy oracle (restatement of qutip/qiskit, oracle/) on a bounded sample, 1 thread
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import workloads  # noqa: E402

METRIC = "wigner_grid_points_per_s_fp64"
UNIT = "points/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--frames", type=int, default=None, help="frames per rank (c3)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sharding", default="auto", choices=["auto", "frames", "grid"],
                    help="N > 1: 'frames' = every GPU its own frames (weak scaling); 'grid' = ONE grid shared out by unit "
                         "range + one reduce (strong scaling; what BASELINE config 5 names). auto: grid for c5, else frames")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl", "none"],
                    help="N>1: how finished grids reach rank 0 (peer = kernels store into rank 0's buffer over NVLink)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--fp64-probe-seconds", type=float, default=1.0, help="length of the DFMA peak probe")
    return ap.parse_args()


def make_workload(name: str, world: int, frames: int | None) -> dict:
    if name == "c3":
        wl = workloads.config3(frames=frames or 64)
    elif name == "c5":
        wl = workloads.config5()
    else:
        wl = workloads.CONFIGS[name]()
    return wl


def describe(wl: dict, world: int) -> dict:
    S = len(wl["xvec"])
    frames = wl["psi"].shape[0] * (len(wl["traced_sets"]) if "traced_sets" in wl else 1)
    return {
        "workload": {
            "c1": "c1: 1 qumode cutoff 16 + 1 qubit, displaced cat alpha=2, 200x200 grid",
            "c2": "c2: 1 qubit + 2 qumodes cutoff 64, cond. displacement + beamsplitter, trace to qumode 0, Wigner 1024x1024",
            "c3": "c3: Jaynes-Cummings frames cutoff 32 + 1 qubit, 400x400 Wigner per frame",
            "c4": "c4: 4 sites cutoff 16, per-site reduced rho (batched partial trace) + Wigner 512x512 each",
            "c5": "c5: single qumode cutoff 1024, Wigner 4096x4096",
        }[wl["name"]],
        "cutoff": int(wl["N"]),
        "grid": [S, S],
        "frames_per_gpu": int(frames),
        "global_frames": int(frames * world),
        "sharding": "frames" if world > 1 else "none",
        "l2": "flushed between timed steps (256 MiB write)",
        "grid_kind": "xvec = yvec = linspace(-6, 6, S) as wigner.py:135,187-188 builds it (mirror-symmetric: the library "
                     "shares one Clenshaw recurrence between the 8 points of equal x^2+p^2; roofline.point_by_point "
                     "times the kernels that do not)",
    }


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock, power and throttle reasons WHILE the timed region runs (NVML, ~2 ms period;
    falls back to `nvidia-smi -lms` when NVML is not importable)."""

    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index: int):
        self.index = index
        self.samples: list[tuple[float, float, int]] = []  # (sm MHz, watts, reason bitmask)
        self.sm_max = None
        self._stop = threading.Event()
        self._thread = None
        self._smi = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def loop():
                while not self._stop.is_set():
                    try:
                        mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        watts = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                        reasons = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        self.samples.append((mhz, watts, reasons))
                    except Exception:
                        pass
                    time.sleep(0.002)

            self._thread = threading.Thread(target=loop, daemon=True)
            self._thread.start()
        except Exception:
            self._start_smi()

    def _start_smi(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.sw_power_cap,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown")
        try:
            self._smi = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

            def pump():
                bits = [0x8, 0x4, 0x20, 0x40]
                for line in self._smi.stdout:
                    parts = [x.strip() for x in line.split(",")]
                    try:
                        mask = sum(b for b, v in zip(bits, parts[3:7]) if v.lower().startswith("active"))
                        self.samples.append((float(parts[0]), float(parts[2]), mask))
                        self.sm_max = float(parts[1])
                    except Exception:
                        continue

            self._thread = threading.Thread(target=pump, daemon=True)
            self._thread.start()
        except Exception:
            self._smi = None

    def stop(self) -> dict:
        self._stop.set()
        if self._smi is not None:
            time.sleep(0.05)
            self._smi.terminate()
        if self._thread is not None:
            self._thread.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["no samples"]}
        sm = [x[0] for x in self.samples]
        power = [x[1] for x in self.samples]
        thr = 0.5 * (min(power) + max(power))
        loaded = [m for m, w, _ in self.samples if w >= thr] or sm  # under load = upper half of the power range
        mask = 0
        for _, w, r in self.samples:
            if w >= thr:
                mask |= r
        reasons = sorted(k for k, b in self.REASONS.items() if mask & b)
        return {"sm_mhz": float(np.median(loaded)), "sm_max_mhz": self.sm_max, "reasons": reasons,
                "power_w_max": float(max(power)), "samples": len(sm), "samples_under_load": len(loaded)}


# ------------------------------------------------------------------------------------------------
def cpu_sample(wl: dict, seconds: float):
    """Times the NumPy oracle on a bounded row sample of the workload; returns (points/s, description)."""
    from oracle import partial_trace_sv, wigner_clenshaw

    psi = wl["psi"][0]
    traced = wl["traced_sets"][0] if "traced_sets" in wl else wl["traced"]
    x = wl["xvec"]
    S = len(x)
    N = wl["N"]
    # NumPy cost model from SURVEY 6: ~1 algorithmic GFLOP/s -> rows that fit in `seconds`
    per_row = S * workloads.flops_per_point(N) / 0.9e9
    rows = int(max(1, min(S, seconds / per_row)))
    idx = np.linspace(0, S - 1, rows + 2).astype(int)[1:-1]  # evenly spaced interior rows
    ys = x[idx]
    t0 = time.perf_counter()
    rho = partial_trace_sv(psi, traced)
    W = wigner_clenshaw(rho, x, ys)
    dt = time.perf_counter() - t0
    assert W.shape == (rows, S)
    return rows * S / dt, f"psi->rho->W on {rows} of {S} grid rows x {S} columns, 1 frame, {dt:.1f} s", rho, idx, W


def _ref_worker(args):
    rho, x, ys = args
    from oracle import wigner_clenshaw

    return wigner_clenshaw(rho, x, ys)


def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's CPU path (NumPy restatement; qutip/qiskit are not installed
    anywhere this runs), fanned out over processes the way animate.py:191 fans frames out."""
    if rank != 0:
        return
    import multiprocessing as mp

    from oracle import partial_trace_sv

    wl = make_workload(args.config, world, args.frames)
    cfg = describe(wl, world)
    psi = wl["psi"][0]
    traced = wl["traced_sets"][0] if "traced_sets" in wl else wl["traced"]
    x = wl["xvec"]
    S, N = len(x), wl["N"]
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    # per step: a row sample sized for ~ (cpu_seconds / steps) of wall time across all workers
    budget = max(1.0, args.cpu_seconds * 4 / max(1, args.steps + args.warmup))
    per_row = S * workloads.flops_per_point(N) / 0.9e9
    rows = int(max(procs, min(S, budget * procs / per_row)))
    rows -= rows % procs
    ys = x[np.linspace(0, S - 1, rows).astype(int)]
    chunks = np.array_split(ys, procs)
    times = []
    with mp.get_context("fork").Pool(procs) as pool:
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            rho = partial_trace_sv(psi, traced)
            pool.map(_ref_worker, [(rho, x, c) for c in chunks])
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = rows * S / (ms * 1e-3)
    sample = f"{rows} of {S} grid rows x {S} columns per step, {procs} worker processes (rows split evenly)"
    # same `scaling` label as the B200 arm of this configuration (config 5 on N > 1 GPUs is one grid: strong)
    one_grid = world > 1 and args.config == "c5" and args.sharding in ("auto", "grid")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if one_grid else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


# ------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def emit(obj) -> None:
    """The ONE JSON line of this run, written to the process's original stdout."""
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    # Libraries (NCCL's version banner, torchrun warnings ...) like to chat on stdout; the contract is one JSON
    # line there, so everything else that is written to fd 1 during the run goes to stderr instead.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from bosonic_qiskit_b200 import get_engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # NCCL prints its version banner on stdout; keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    eng = get_engine(local)
    wl = make_workload(args.config, world, args.frames)
    cfg = describe(wl, world)
    x_np = wl["xvec"]
    S, N = len(x_np), int(wl["N"])
    multi_sets = "traced_sets" in wl
    frames = cfg["frames_per_gpu"]
    points_rank = frames * S * S
    # one grid on all GPUs (BASELINE config 5): every rank evaluates an equal range of the grid's units, one reduce
    one_grid = world > 1 and not multi_sets and frames == 1 and (args.sharding == "grid" or (args.sharding == "auto" and args.config == "c5"))
    if one_grid:
        if not eng.grid_is_mirror(x_np):
            raise SystemExit("--sharding grid needs the mirror-symmetric grid the reference builds")
        from bosonic_qiskit_b200.parallel import shard_range

        u_first, u_last = shard_range(eng.grid_units(S), rank, world)
        cfg.update({"frames_per_gpu": 1.0 / world, "global_frames": 1,
                    "sharding": "one grid: equal ranges of its units (pairs of grid lines, 8 points each) per GPU, then one NCCL reduce to rank 0"})
    points_global = points_rank if one_grid else points_rank * world
    points_launch = points_rank / world if one_grid else points_rank

    psi_t = torch.from_numpy(wl["psi"]).to(dev)
    x_t = torch.from_numpy(x_np).to(dev)
    W_t = torch.empty((frames, S, S), dtype=torch.float64, device=dev)
    rho_t = torch.empty((frames, N, N), dtype=torch.complex128, device=dev)
    gathered = None
    peer = None
    gather_mode = "reduce" if one_grid else args.gather if (world > 1 and not multi_sets) else ("nccl" if world > 1 and args.gather != "none" else "none")
    if world > 1 and gather_mode == "nccl":
        gathered = torch.empty((world * frames, S, S), dtype=torch.float64, device=dev) if rank == 0 else None
    if world > 1 and gather_mode == "peer":
        from bosonic_qiskit_b200.parallel import PeerGatherBuffer

        peer = PeerGatherBuffer(eng, (world * frames, S, S), dst=0)  # collective; outside the timed region
        peer_out = peer.ptr_for(rank * frames)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()

    def step_device():
        if multi_sets:
            # c4: several traces of one psi, then one batched Wigner launch
            import ctypes

            from bosonic_qiskit_b200._lib import check
            sets = wl["traced_sets"]
            flat = (ctypes.c_int * (len(sets) * len(sets[0])))(*[q for s in sets for q in s])
            check(eng._lib.bq_partial_trace_sv_multi_dev(eng._h, ctypes.c_void_p(psi_t.data_ptr()), wl["n_qubits"], flat,
                                                         len(sets[0]), len(sets), ctypes.c_void_p(rho_t.data_ptr()),
                                                         ctypes.c_void_p(stream.cuda_stream)))
            eng.wigner_t(rho_t, x_t, out=W_t)
        elif one_grid:
            eng.partial_trace_sv_t(psi_t, wl["traced"], out=rho_t)  # 16 KiB of psi: replicated on every rank
            W_t.zero_()
            eng.wigner_t(rho_t, x_t, out=W_t, unit_range=(u_first, u_last - u_first))
            dist.reduce(W_t, dst=0, op=dist.ReduceOp.SUM)  # the shares touch disjoint points: the sum assembles the grid
        elif peer is not None:
            # fused compute + gather: the kernel's stores land in rank 0's buffer (NVLink peer memory)
            eng.state_to_wigner_t(psi_t, wl["traced"], x_t, out=peer_out, rho_out=rho_t)
        else:
            eng.state_to_wigner_t(psi_t, wl["traced"], x_t, out=W_t, rho_out=rho_t)
        if world > 1 and gather_mode == "nccl":
            if rank == 0:
                dist.gather(W_t, list(gathered.view(world, frames, S, S).unbind(0)), dst=0)
            else:
                dist.gather(W_t, None, dst=0)

    # ---- device-resident timing ------------------------------------------------------------
    step_device()  # (first call: tables, workspaces, kernel attributes)
    torch.cuda.synchronize()
    fp64_peak = eng.fp64_peak_tflops(args.fp64_probe_seconds)  # sustained DFMA rate on this GPU, right before the timed region
    # the untimed warm-up steps come AFTER the probe, directly before the timed region, with the same L2 flush between
    # them and with the per-launch event bracket already on (the first steps after the 1 s probe run ~10 % slow: clocks
    # and the host-side launch path settle) -- at least 10 of them, more if asked for
    n_warm = max(args.warmup, 10)
    eng.profile(True)
    for _ in range(n_warm):
        flush.fill_(1.0)
        step_device()
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.profile(True)
    eng.profile_read()
    launches0 = eng.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for _ in range(8):
        flush.fill_(1.0)  # ~0.3 ms of queued device work: the host is ahead of the GPU from the first timed step on
    for a, b in ev:
        flush.fill_(1.0)  # L2 flush, outside the per-step events
        a.record()
        step_device()
        b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = eng.launches - launches0
    kernel_name = eng.last_wigner_kernel()
    kern_ms, kern_n = eng.profile_read()
    eng.profile(False)
    total_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = points_global / (ms_per_step * 1e-3)

    # ---- the same step with recurrence sharing off: the point-by-point kernels (the plain fp64-roofline case) ------
    pbp = None
    shared_used = bool(eng.grid_is_mirror(x_np))
    if world == 1 and shared_used:
        W_keep = W_t.clone()
        eng.set_grid_sharing(False)
        try:
            for _ in range(3):
                step_device()
            torch.cuda.synchronize()
            eng.profile(True)
            eng.profile_read()
            pev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            for a, b in pev:
                flush.fill_(1.0)
                a.record()
                step_device()
                b.record()
            torch.cuda.synchronize()
            p_ms, p_n = eng.profile_read()
            eng.profile(False)
            pbp = {"kernel_ms": p_ms / max(p_n, 1), "launches": p_n, "kernel": eng.last_wigner_kernel(),
                   "ms_per_step": float(sum(a.elapsed_time(b) for a, b in pev)) / args.steps,
                   "max_abs_diff_to_shared_over_max_W": float((W_t - W_keep).abs().max().item() / W_keep.abs().max().item())}
        finally:
            eng.set_grid_sharing(True)
        W_t.copy_(W_keep)

    # ---- accuracy of what was just timed (rank 0, a row sample) ---------------------------------
    # ---- end-to-end through the public NumPy API (host buffers) ---------------------------------
    e2e = None
    if one_grid:
        from bosonic_qiskit_b200.parallel import wigner_row_slabs

        psi_pin = eng.pinned_empty(wl["psi"].shape, np.complex128)
        psi_pin[...] = wl["psi"]
        out_pin = torch.empty((S, S), dtype=torch.float64, pin_memory=True) if rank == 0 else None

        def step_host():
            rho_h = eng.partial_trace_sv(psi_pin[0], wl["traced"])
            Wg, _ = wigner_row_slabs(rho_h, x_np, engine=eng)
            if rank == 0:
                out_pin.copy_(Wg)  # device -> pinned host (synchronous)
            dist.barrier()

        step_host()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        e2e = {"value": points_global * args.steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(psi_pin.nbytes + world * N * N * 16), "d2h_bytes_per_step": int(S * S * 8 + N * N * 16),
               "ms_per_step": 1e3 * dt / args.steps,
               "api": "Engine.partial_trace_sv + parallel.wigner_row_slabs (rho to every GPU, unit ranges, reduce) + copy to host on rank 0"}
    elif not multi_sets:
        psi_pin = eng.pinned_empty(wl["psi"].shape, np.complex128)
        psi_pin[...] = wl["psi"]
        out_pin = eng.pinned_empty((frames, S, S), np.float64)
        for _ in range(3):
            eng.state_to_wigner(psi_pin, wl["traced"], x_np, out=out_pin)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.state_to_wigner(psi_pin, wl["traced"], x_np, out=out_pin)
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": points_rank * world * args.steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(psi_pin.nbytes), "d2h_bytes_per_step": int(out_pin.nbytes),
               "ms_per_step": 1e3 * dt / args.steps, "api": "Engine.state_to_wigner -> bq_state_to_wigner_host"}
    else:
        import bosonic_qiskit_b200  # noqa: F401

        psi_np = wl["psi"][0]
        for _ in range(3):
            rhos = eng.partial_trace_sv_multi(psi_np, wl["traced_sets"])
            eng.wigner(rhos, x_np)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            rhos = eng.partial_trace_sv_multi(psi_np, wl["traced_sets"])
            Wh = eng.wigner(rhos, x_np)
        dt = time.perf_counter() - t0
        e2e = {"value": points_rank * world * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(psi_np.nbytes + rhos.nbytes),
               "d2h_bytes_per_step": int(Wh.nbytes + rhos.nbytes), "ms_per_step": 1e3 * dt / args.steps,
               "api": "Engine.partial_trace_sv_multi + Engine.wigner (host entry points)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"

    traffic = None  # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get(wl["name"])
        if t:
            traffic = {"bytes": t["dram_bytes_read"] + t["dram_bytes_write"], "source": t["source"]}
    except Exception:
        pass
    F = workloads.flops_per_point(N)
    kern_avg_ms = kern_ms / max(kern_n, 1)
    achieved_tf = points_launch * F / (kern_avg_ms * 1e-3) / 1e12
    big = points_rank >= 148 * 2048
    if shared_used:
        kname = "wigner_sym_resident_kernel" if N <= 64 else ("wigner_sym_diag_kernel" if big else "wigner_stream_kernel")
    else:
        kname = "wigner_resident_kernel" if N <= 64 else ("wigner_diag_kernel" if big else "wigner_stream_kernel")
    kname = kernel_name
    roofline = {
        "bound": "fp64", "kernel": kname,
        "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak,
        "peak_source": "DFMA probe kernel run for 1 s on this GPU immediately before the timed region (bq_fp64_peak); "
                       "nominal 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2",
        "flops_per_point": F, "points_per_launch": points_launch, "kernel_ms": kern_avg_ms, "kernel_launches_timed": kern_n,
        "kernel_share_of_step": kern_avg_ms / ms_per_step,
        "traffic": traffic,
        "hbm_store": {"bound": "hbm", "achieved": points_launch * 8 / (kern_avg_ms * 1e-3) / 1e9, "peak": hbm_peak,
                      "unit": "GB/s", "peak_source": hbm_src,
                      "frac": points_launch * 8 / (kern_avg_ms * 1e-3) / 1e9 / hbm_peak},
    }
    if shared_used:
        # `achieved` above is the ALGORITHMIC work of the reference's point-by-point recurrence over the
        # kernel's time: it exceeds the fp64 peak because the kernel shares one recurrence between 8 grid points.
        # What the fp64 pipe actually executed, and the point-by-point kernels timed in the same run, are next to it.
        units = (u_last - u_first) if one_grid else frames * workloads.shared_units(S)
        ex_tf = units * workloads.flops_per_unit_shared(N) / (kern_avg_ms * 1e-3) / 1e12
        roofline["note"] = ("frac > 1: algorithmic flops of the reference recurrence (one per grid point) / kernel time; the "
                            "symmetric-grid kernel runs one recurrence per pair of grid lines. executed = flops it really "
                            "issues; point_by_point = the unshared kernels, same inputs, same run")
        roofline["executed"] = {"flops_per_launch": units * workloads.flops_per_unit_shared(N), "units_per_launch": units,
                                "achieved": ex_tf, "frac": ex_tf / fp64_peak}
        if pbp is not None:
            p_tf = points_rank * F / (pbp["kernel_ms"] * 1e-3) / 1e12
            roofline["point_by_point"] = {
                "kernel": pbp["kernel"], "kernel_ms": pbp["kernel_ms"], "achieved": p_tf, "frac": p_tf / fp64_peak, "ms_per_step": pbp["ms_per_step"],
                "value": points_rank / (pbp["ms_per_step"] * 1e-3),
                "max_abs_diff_to_shared_over_max_W": pbp["max_abs_diff_to_shared_over_max_W"]}

    # parity of the timed output against the oracle (row sample) and the CPU baseline, same inputs
    cpu = None
    parity = None
    if not args.no_cpu_baseline and world == 1:
        v, sample, rho_ref, idx, W_ref = cpu_sample(wl, args.cpu_seconds)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": sample + " (NumPy restatement of qiskit.partial_trace + qutip.wigner clenshaw; "
                                  "qutip/qiskit are not installed)"}
        W_dev = W_t[0].cpu().numpy()[idx]
        rho_dev = rho_t[0].cpu().numpy()
        parity = {"max_abs_dW_over_max_W": float(np.abs(W_dev - W_ref).max() / float(W_t[0].abs().max().item())),
                  "max_abs_drho_over_max_rho": float(np.abs(rho_dev - rho_ref).max() / np.abs(rho_ref).max()),
                  "tolerance": 1e-10}

    if one_grid:
        # the assembled grid against rows evaluated on this GPU alone by the point-by-point kernels
        rows = torch.arange(0, S, max(S // 16, 1), device=dev)
        W_rows = eng.wigner_t(rho_t[0], x_t, x_t[rows].contiguous())
        parity = {"max_abs_dW_over_max_W_vs_rows_on_one_gpu": float((W_t[0][rows] - W_rows).abs().max().item() / W_t[0].abs().max().item()),
                  "rows_checked": int(rows.numel()), "tolerance": 1e-10}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": n_warm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if one_grid else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": cfg, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
        "gather": None if world == 1 else {"peer": "kernel stores finished tiles into rank 0's buffer over NVLink (CUDA IPC peer memory), inside the step",
                                             "nccl": "torch.distributed.gather (NCCL) to rank 0 inside the step",
                                             "reduce": "torch.distributed.reduce (NCCL, sum of disjoint shares) to rank 0 inside the step",
                                             "none": "none (results stay sharded)"}[gather_mode],
    }
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
