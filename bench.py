#!/usr/bin/env python
"""Benchmark of the statevector -> reduced rho -> Wigner grid hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c5] [--legs c1,c2,c3,c4,k1]

One "step" is one pass of the hot path over one batch of synthetic input.  The headline workload is BASELINE.json
configs[4] ("c5": single qumode at cutoff 1024, Wigner 4096 x 4096) -- the only named configuration at cutoff >= 256,
the one north_star states its fp64-roofline target on:  psi (1024 amplitudes) -> rho (1024 x 1024) -> W (4096 x 4096).

  N = 1      the whole grid on one GPU.
  N > 1      (torchrun, one rank per GPU) ONE grid shared out over the GPUs by equal ranges of its units, assembled on
             rank 0 by one NCCL reduce inside the step: "scaling": "strong".

  value      whole-job Wigner grid points / s, inputs resident in HBM (CUDA events per step, max over ranks)
  e2e        the same metric through the public host API: pinned host psi -> H2D -> kernels -> D2H -> ONE host array.
             N = 1: Engine.state_to_wigner (bq_state_to_wigner_host); N > 1: MultiEngine (bq_state_to_wigner_sharded_host),
             ONE process driving all N GPUs, called by rank 0 while the other ranks idle on a CPU barrier
  roofline   dominant kernel: fp64 flops it EXECUTES / its own event time against an fp64-FMA peak measured on this GPU
             in the same run (frac <= 1); `algorithmic` = the reference recurrence's flops per grid point over
             the same time -- above the peak on the reference's grids, because the kernel shares one recurrence between
             the 8 points of equal x^2 + p^2; `point_by_point` = the kernels that do not, same inputs, same run
  cpu_baseline / --impl reference   the reference's CPU path on the box's host cores: qutip + qiskit themselves when
             they import ("kind": "reference"), otherwise the NumPy restatement under oracle/ ("kind": "port")
  parity     rank 0, every N: the timed output against the CPU path on a sub-grid, max|dW| / max|W| (tolerance 1e-10)
  legs       the other named configurations (c1, c2, c3 with 512 frames, c4) and a partial trace large enough for the
             FP64 tensor-core kernel (k1), each timed the same way, inside the same JSON line
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
WORKLOAD_TEXT = {
    "c1": "c1: 1 qumode cutoff 16 + 1 qubit, displaced cat alpha=2, 200x200 grid",
    "c2": "c2: 1 qubit + 2 qumodes cutoff 64, cond. displacement + beamsplitter, trace to qumode 0, Wigner 1024x1024",
    "c3": "c3: Jaynes-Cummings frames cutoff 32 + 1 qubit, 400x400 Wigner per frame",
    "c4": "c4: 4 sites cutoff 16, per-site reduced rho (batched partial trace) + Wigner 512x512 each",
    "c5": "c5: single qumode cutoff 1024, Wigner 4096x4096",
}
GRID_KIND = ("xvec = yvec = linspace(-6, 6, S) as wigner.py:135,187-188 builds it (mirror-symmetric: the library shares one "
             "Clenshaw recurrence between the 8 points of equal x^2+p^2; roofline.point_by_point times the kernels that do not)")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c5", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--frames", type=int, default=None, help="frames of the c3 headline workload (all GPUs together)")
    ap.add_argument("--legs", default="c1,c2,c3,c4,k1", help="other workloads timed into the same JSON line ('' = none)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-point-by-point", action="store_true")
    ap.add_argument("--sharding", default="auto", choices=["auto", "frames", "grid"],
                    help="N > 1: 'grid' = ONE grid shared out by unit range + one reduce (strong scaling; what BASELINE "
                         "config 5 names); 'frames' = the frames of a batch split over the GPUs (strong). auto: by workload")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--fp64-probe-seconds", type=float, default=1.0, help="length of the DFMA peak probe")
    return ap.parse_args()


def make_workload(name: str, frames: int | None = None) -> dict:
    if name == "c3":
        return workloads.config3(frames=frames or 512)
    return workloads.CONFIGS[name]()


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
# the reference's CPU path: qutip + qiskit themselves when they import, else the restatement under oracle/
# ------------------------------------------------------------------------------------------------
def cpu_path():
    """(kind, partial_trace(psi, traced) -> rho, wigner(rho, xvec, yvec) -> W[iy, ix])."""
    try:
        import qiskit.quantum_info as qi
        import qutip

        def ptrace(psi, traced):
            if not traced:
                return np.asarray(qi.DensityMatrix(qi.Statevector(psi)).data)
            return np.asarray(qi.partial_trace(qi.Statevector(psi), list(traced)).data)  # util.py:580,596,619,693

        def wig(rho, xvec, yvec):
            return np.asarray(qutip.wigner(qutip.Qobj(rho), xvec, yvec, g=np.sqrt(2), method="clenshaw"))  # wigner.py:190

        return "reference", ptrace, wig, f"qutip {qutip.__version__} + qiskit {__import__('qiskit').__version__}"
    except Exception:
        from oracle import partial_trace_sv, wigner_clenshaw

        return "port", partial_trace_sv, wigner_clenshaw, ("NumPy restatement of qiskit.partial_trace + qutip.wigner clenshaw "
                                                          "(oracle/); qutip / qiskit do not import here")


def sub_grid(S: int, N: int, seconds: float, rate: float = 0.9e9):
    """Row and column indices of an evenly spaced sub-grid the CPU path evaluates in about `seconds`
    (NumPy cost model of SURVEY 6: ~0.9 algorithmic GFLOP/s per core)."""
    pts = max(16.0, seconds * rate / workloads.flops_per_point(N))
    if pts >= S * S:
        return np.arange(S), np.arange(S)
    cols = S if pts >= S else max(8, int(pts ** 0.5) * 2)
    rows = int(max(1, min(S, pts // cols)))
    iy = np.linspace(0, S - 1, rows + 2).astype(int)[1:-1] if rows < S else np.arange(S)
    ix = np.linspace(0, S - 1, cols).astype(int) if cols < S else np.arange(S)
    return np.unique(iy), np.unique(ix)


def cpu_sample(wl: dict, seconds: float):
    """Times the CPU path (1 thread) on a bounded sub-grid of the workload's first frame."""
    kind, ptrace, wig, what = cpu_path()
    psi = wl["psi"][0]
    traced = wl["traced_sets"][0] if "traced_sets" in wl else wl["traced"]
    x = wl["xvec"]
    S, N = len(x), wl["N"]
    iy, ix = sub_grid(S, N, seconds)
    t0 = time.perf_counter()
    rho = ptrace(psi, traced)
    W = wig(rho, x[ix], x[iy])
    dt = time.perf_counter() - t0
    assert W.shape == (len(iy), len(ix))
    pts = len(iy) * len(ix)
    sample = f"psi->rho->W on a {len(iy)} x {len(ix)} sub-grid ({pts} of {S * S} points) of 1 frame, {dt:.1f} s; {what}"
    return dict(value=pts / dt, kind=kind, sample=sample, rho=rho, iy=iy, ix=ix, W=W)


def _ref_worker(args):
    rho, xs, ys = args
    _, _, wig, _ = cpu_path()
    return wig(rho, xs, ys)


def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's CPU path on all host cores, fanned out over processes the way animate.py:191
    fans frames out (rank 0 only; each step a bounded sub-grid of the headline workload)."""
    if rank != 0:
        return
    import multiprocessing as mp

    kind, ptrace, _, what = cpu_path()
    wl = make_workload(args.config, args.frames)
    cfg = describe(wl, world, args)
    psi = wl["psi"][0]
    traced = wl["traced_sets"][0] if "traced_sets" in wl else wl["traced"]
    x = wl["xvec"]
    S, N = len(x), wl["N"]
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    # per step: every worker one row segment, sized so that warm-up + steps take about 4 x cpu_seconds of wall time
    budget = max(0.5, args.cpu_seconds * 4 / max(1, args.steps + args.warmup))
    # (at least 1024 points per worker call: below that NumPy's per-operation overhead, not its arithmetic, is what is timed)
    pts_per_proc = max(1024, int(budget * 0.9e9 / workloads.flops_per_point(N)))
    cols = min(S, pts_per_proc)
    rows_per_proc = max(1, min(S // procs, pts_per_proc // cols))
    ix = np.linspace(0, S - 1, cols).astype(int) if cols < S else np.arange(S)
    iy = np.linspace(0, S - 1, rows_per_proc * procs).astype(int)
    chunks = np.array_split(x[iy], procs)
    times = []
    with mp.get_context("fork").Pool(procs) as pool:
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            rho = ptrace(psi, traced)
            pool.map(_ref_worker, [(rho, x[ix], c) for c in chunks])
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
    ms = 1e3 * float(np.mean(times))
    pts = len(iy) * len(ix)
    value = pts / (ms * 1e-3)
    sample = (f"{len(iy)} x {len(ix)} sub-grid ({pts} of {S * S} points) of 1 frame per step, {procs} worker processes "
              f"(rows split evenly); {what}")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": cfg.pop("_scaling"), "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


def describe(wl: dict, world: int, args) -> dict:
    """The `config` object (same for both arms); `_scaling` is popped by the caller."""
    S = len(wl["xvec"])
    multi_sets = "traced_sets" in wl
    frames = wl["psi"].shape[0] * (len(wl["traced_sets"]) if multi_sets else 1)
    cfg = {"workload": WORKLOAD_TEXT[wl["name"]], "cutoff": int(wl["N"]), "grid": [S, S], "global_frames": int(frames),
           "l2": "flushed between timed steps (256 MiB write)", "grid_kind": GRID_KIND}
    mode = sharding_mode(wl, world, args)
    if mode == "grid":
        cfg.update({"frames_per_gpu": 1.0 / world,
                    "sharding": "one grid: equal ranges of its units (pairs of grid lines, 8 points each) per GPU, then one NCCL reduce to rank 0"})
    elif mode == "frames":
        cfg.update({"frames_per_gpu": frames / world, "sharding": "contiguous frame blocks per GPU, results stay on the GPU that made them"})
    elif mode == "replicas":
        cfg.update({"frames_per_gpu": frames, "global_frames": int(frames * world),
                    "sharding": "replicas: every GPU the whole workload (a single frame cannot be split by frame)"})
    else:
        cfg.update({"frames_per_gpu": frames, "sharding": "none"})
    # (the label says how the work per GPU changes with N for this workload, so it is the same at N = 1 as at N = 8)
    cfg["_scaling"] = "weak" if sharding_mode(wl, max(world, 2), args) == "replicas" else "strong"
    return cfg


def sharding_mode(wl: dict, world: int, args) -> str:
    if world == 1:
        return "none"
    frames = wl["psi"].shape[0]
    if "traced_sets" in wl:
        return "replicas"
    if frames == 1:
        return "grid" if args.sharding in ("auto", "grid") else "replicas"
    return "frames" if args.sharding in ("auto", "frames") else "replicas"


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


class Bench:
    """Device handles shared by the headline measurement and the legs."""

    def __init__(self, args, rank, world, local):
        import torch
        import torch.distributed as dist

        from bosonic_qiskit_b200 import get_engine

        self.torch, self.dist = torch, dist
        self.args, self.rank, self.world, self.local = args, rank, world, local
        self.dev = torch.device("cuda", local)
        self.eng = get_engine(local)
        self.flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=self.dev)
        self.cpu_group = None
        if world > 1:
            self.cpu_group = dist.new_group(backend="gloo")  # host-side barriers: waiting ranks leave their GPUs idle
        self.fp64_peak = None
        self.multi = None

    # ---- timing ---------------------------------------------------------------------------------------
    def time_steps(self, step, steps: int, warm: int, profile: bool = True):
        """`warm` untimed steps, barrier + synchronize, then exactly `steps` steps, each between its own pair of CUDA
        events with an L2 flush before it; returns (ms per step = max over ranks, kernel ms per launch, kernel launches
        per step, library launches per step)."""
        torch, dist, eng = self.torch, self.dist, self.eng
        for _ in range(warm):
            self.flush.fill_(1.0)
            step()
        torch.cuda.synchronize()
        if profile:
            eng.profile(True)
            eng.profile_read()
        launches0 = eng.launches
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        for _ in range(8):
            self.flush.fill_(1.0)  # ~0.3 ms of queued device work: the host is ahead of the GPU from the first timed step on
        for a, b in ev:
            self.flush.fill_(1.0)  # L2 flush, outside the per-step events
            a.record()
            step()
            b.record()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        launches = eng.launches - launches0
        kern_ms, kern_n = eng.profile_read() if profile else (0.0, 0)
        if profile:
            eng.profile(False)
        total_ms = float(sum(a.elapsed_time(b) for a, b in ev))
        if self.world > 1:
            t = torch.tensor([total_ms], dtype=torch.float64, device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms / steps, (kern_ms / kern_n if kern_n else None), kern_n / steps, launches / steps

    def host_barrier(self):
        if self.cpu_group is not None:
            self.dist.barrier(group=self.cpu_group)

    def multi_engine(self):
        """rank 0: ONE process driving all `world` GPUs (bq_multi_*); the other ranks never create it."""
        if self.multi is None:
            from bosonic_qiskit_b200 import MultiEngine

            self.multi = MultiEngine(list(range(self.world)))
        return self.multi

    def time_host(self, call, steps: int, warm: int = 2):
        """Wall clock of `call` (a public host-API call that returns with its result in host memory).  N > 1: only rank 0
        calls -- through the single-process multi-GPU engine -- while the other ranks sit on a CPU barrier."""
        if self.world > 1 and self.rank != 0:
            self.host_barrier()
            self.host_barrier()
            return None
        for _ in range(warm):
            call()
        self.host_barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            call()
        dt = time.perf_counter() - t0
        self.host_barrier()
        return dt / steps

    # ---- one frame-batch workload (c1 .. c5 on one GPU, or a rank's share of it) -------------------------
    def frames_workload(self, wl: dict, steps: int, warm: int, mode: str, pbp_steps: int = 0, want_cpu: bool = False,
                        cpu_seconds: float = 0.0, e2e_steps: int | None = None):
        torch, dist, eng, dev, world, rank = self.torch, self.dist, self.eng, self.dev, self.world, self.rank
        x_np = wl["xvec"]
        S, N = len(x_np), int(wl["N"])
        multi_sets = "traced_sets" in wl
        psi_all = wl["psi"]
        n_sets = len(wl["traced_sets"]) if multi_sets else 1
        frames_all = psi_all.shape[0] * n_sets
        from bosonic_qiskit_b200.parallel import shard_range

        lo, hi = (0, psi_all.shape[0])
        if mode == "frames":
            lo, hi = shard_range(psi_all.shape[0], rank, world)
        psi_np = np.ascontiguousarray(psi_all[lo:hi])
        frames = (hi - lo) * n_sets
        mirror = bool(eng.grid_is_mirror(x_np))
        u_first = u_count = None
        if mode == "grid":
            if not mirror:
                raise SystemExit("--sharding grid needs the mirror-symmetric grid the reference builds")
            a, b = shard_range(eng.grid_units(S), rank, world)
            u_first, u_count = a, b - a
        psi_t = torch.from_numpy(psi_np).to(dev) if frames else None
        x_t = torch.from_numpy(x_np).to(dev)
        W_t = torch.empty((max(frames, 1), S, S), dtype=torch.float64, device=dev)
        rho_t = torch.empty((max(frames, 1), N, N), dtype=torch.complex128, device=dev)

        def step():
            if frames == 0:
                return
            if multi_sets:
                eng.state_to_wigner_multi_t(psi_t[0], wl["traced_sets"], x_t, out=W_t, rho_out=rho_t)
            elif mode == "grid":
                eng.partial_trace_sv_t(psi_t, wl["traced"], out=rho_t)  # 16 KiB of psi: replicated on every rank
                W_t.zero_()
                eng.wigner_t(rho_t, x_t, out=W_t, unit_range=(u_first, u_count))
                dist.reduce(W_t, dst=0, op=dist.ReduceOp.SUM)  # the shares touch disjoint points: the sum assembles the grid
            else:
                eng.state_to_wigner_t(psi_t, wl["traced"], x_t, out=W_t, rho_out=rho_t)

        step()  # first call: tables, workspaces, kernel attributes
        torch.cuda.synchronize()
        if self.fp64_peak is None:
            self.fp64_peak = eng.fp64_peak_tflops(self.args.fp64_probe_seconds)  # DFMA rate of this GPU, before the timed region
            if world > 1:  # one denominator for the job: the highest rate any of its (identical) GPUs reached
                t = torch.tensor([self.fp64_peak], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                self.fp64_peak = float(t.item())
        ms, kern_ms, kern_per_step, launches = self.time_steps(step, steps, warm)
        kernel_name = eng.last_wigner_kernel()
        points_global = frames_all * S * S * (world if mode == "replicas" else 1)
        points_launch = (S * S / world) if mode == "grid" else frames * S * S
        res = {"value": points_global / (ms * 1e-3), "ms_per_step": ms, "gpu_launches_per_step": launches,
               "frames": frames_all if mode != "replicas" else frames_all * world, "cutoff": N, "grid": [S, S]}

        # ---- roofline of the dominant (Wigner) kernel ----
        F = workloads.flops_per_point(N)
        peak = self.fp64_peak
        roof = None
        if kern_ms:
            kern_step_ms = kern_ms * kern_per_step  # a step may split its Wigner work over two launches
            alg_tf = points_launch * F / (kern_step_ms * 1e-3) / 1e12
            roof = {"bound": "fp64", "kernel": kernel_name, "unit": "TFLOP/s", "peak": peak,
                    "peak_source": "DFMA probe kernel run for 1 s on this GPU before the timed region (bq_fp64_peak; "
                                   "MEASURED_PEAKS.json has no fp64 entry); nominal 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2",
                    "kernel_ms": kern_step_ms, "kernel_launches_per_step": kern_per_step, "kernel_share_of_step": kern_step_ms / ms,
                    "points_per_launch": points_launch,
                    "algorithmic": {"flops_per_point": F, "achieved": alg_tf, "frac": alg_tf / peak}}
            if mirror and kernel_name.startswith("wigner_sym"):
                units = u_count if mode == "grid" else frames * workloads.shared_units(S)
                ex_tf = units * workloads.flops_per_unit_shared(N) / (kern_step_ms * 1e-3) / 1e12
                roof.update({"achieved": ex_tf, "frac": ex_tf / peak, "flops_per_unit_executed": workloads.flops_per_unit_shared(N),
                             "units_per_launch": units,
                             "note": "achieved / frac: the fp64 flops the kernel EXECUTES (one recurrence per unit = pair of grid lines "
                                     "= 8 grid points); algorithmic.frac divides the reference's per-point flops by the same time and "
                                     "exceeds 1 because of that sharing"})
            else:
                roof.update({"achieved": alg_tf, "frac": alg_tf / peak})
            roof["hbm_store"] = {"bound": "hbm", "achieved": points_launch * 8 / (kern_step_ms * 1e-3) / 1e9, "unit": "GB/s"}
        res["roofline"] = roof

        # ---- the same step with recurrence sharing off: the point-by-point kernels ----
        if pbp_steps and mirror and world == 1 and roof and not multi_sets:
            W_keep = W_t.clone()
            eng.set_grid_sharing(False)
            try:
                step()
                torch.cuda.synchronize()
                p_ms, p_kern, p_per, _ = self.time_steps(step, pbp_steps, 1)
                diff = float((W_t - W_keep).abs().max().item() / W_keep.abs().max().item())
                p_tf = points_launch * F / (p_kern * p_per * 1e-3) / 1e12
                roof["point_by_point"] = {"kernel": eng.last_wigner_kernel(), "kernel_ms": p_kern * p_per, "achieved": p_tf,
                                          "frac": p_tf / peak, "ms_per_step": p_ms, "steps": pbp_steps,
                                          "value": points_launch / (p_ms * 1e-3), "max_abs_diff_to_shared_over_max_W": diff}
            finally:
                eng.set_grid_sharing(True)
            W_t.copy_(W_keep)
            del W_keep

        # ---- end to end through the public host API: host psi in, ONE host array out ----
        e2e_n = e2e_steps if e2e_steps is not None else steps
        if world == 1:
            psi_pin = eng.pinned_empty(psi_all.shape, np.complex128)
            psi_pin[...] = psi_all
            if multi_sets:
                out_pin = eng.pinned_empty((frames_all, S, S), np.float64)
                call = lambda: eng.state_to_wigner_multi(psi_pin[0], wl["traced_sets"], x_np, out=out_pin)  # noqa: E731
                api = "Engine.state_to_wigner_multi -> bq_state_to_wigner_multi_host"
            else:
                out_pin = eng.pinned_empty((frames_all, S, S), np.float64)
                call = lambda: eng.state_to_wigner(psi_pin, wl["traced"], x_np, out=out_pin)  # noqa: E731
                api = "Engine.state_to_wigner -> bq_state_to_wigner_host"
            dt = self.time_host(call, e2e_n)
            h2d, d2h, pts = psi_pin.nbytes, out_pin.nbytes, frames_all * S * S
        else:
            dt, h2d, d2h, api, pts = None, 0, 0, "", frames_all * S * S
            out_pin = None
            has_call = not (multi_sets or mode == "replicas")  # (same verdict on every rank: time_host is collective)
            call = None
            if rank == 0 and has_call:
                me = self.multi_engine()
                out_pin = me.pinned_empty((frames_all, S, S), np.float64)
                psi_pin = me.pinned_empty(psi_all.shape, np.complex128)
                psi_pin[...] = psi_all
                if mode == "grid":
                    call = lambda: me.state_to_wigner(psi_pin, wl["traced"], x_np, shard="grid", out=out_pin)  # noqa: E731
                    api = ("MultiEngine.state_to_wigner(shard='grid') -> bq_state_to_wigner_sharded_host: ONE process, psi to every GPU, "
                           "unit ranges stored into one frame on GPU 0 over NVLink peer memory, one copy to the host")
                    h2d, d2h = psi_pin.nbytes * world, out_pin.nbytes
                else:
                    call = lambda: me.state_to_wigner(psi_pin, wl["traced"], x_np, shard="frames", out=out_pin)  # noqa: E731
                    api = ("MultiEngine.state_to_wigner(shard='frames') -> bq_state_to_wigner_sharded_host: ONE process, frame blocks per "
                           "GPU, every GPU copies its grids into its slice of the one pinned host array over its own PCIe link")
                    h2d, d2h = psi_pin.nbytes, out_pin.nbytes
            if has_call:
                torch.cuda.synchronize()
                dt = self.time_host(call, e2e_n)
        res["e2e"] = None
        if dt is not None:
            res["e2e"] = {"value": pts / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                          "ms_per_step": 1e3 * dt, "steps": e2e_n, "api": api}

        # ---- parity of the timed output (and of the host API's output) against the CPU path, rank 0 ----
        if want_cpu and rank == 0:
            cs = cpu_sample(wl, cpu_seconds)
            W0 = W_t[0]
            idx_y = torch.from_numpy(cs["iy"]).to(dev)
            idx_x = torch.from_numpy(cs["ix"]).to(dev)
            W_dev = W0.index_select(0, idx_y).index_select(1, idx_x).cpu().numpy()
            wmax = float(W0.abs().max().item())
            par = {"max_abs_dW_over_max_W": float(np.abs(W_dev - cs["W"]).max() / wmax), "points_checked": int(W_dev.size),
                   "against": cs["kind"], "tolerance": 1e-10}
            if frames:
                rho_dev = rho_t[0].cpu().numpy()
                par["max_abs_drho_over_max_rho"] = float(np.abs(rho_dev - cs["rho"]).max() / np.abs(cs["rho"]).max())
            if res["e2e"] is not None and out_pin is not None:
                W_host = np.asarray(out_pin[0])[np.ix_(cs["iy"], cs["ix"])]
                par["e2e_max_abs_dW_over_max_W"] = float(np.abs(W_host - cs["W"]).max() / wmax)
                par["e2e_equals_device_result"] = bool(np.array_equal(np.asarray(out_pin[0]), W0.cpu().numpy()))
            res["parity"] = par
            res["cpu_baseline"] = {"value": cs["value"], "unit": UNIT, "cores": 1, "kind": cs["kind"], "sample": cs["sample"]}
        del W_t, rho_t
        torch.cuda.empty_cache()
        return res

    # ---- K1 where it is not launch-bound: rho = Psi Psi^dagger, Psi 4096 x 8192 (keep two cutoff-64 modes) -------------
    def k1_leg(self, steps: int = 5):
        torch, eng, dev = self.torch, self.eng, self.dev
        n_qubits, n_keep = 25, 12
        traced = list(range(n_keep, n_qubits))
        g = torch.Generator(device=dev).manual_seed(7)
        psi = torch.randn((1 << n_qubits, 2), dtype=torch.float64, device=dev, generator=g)
        psi = torch.view_as_complex(psi / psi.norm())
        D, T = 1 << n_keep, 1 << (n_qubits - n_keep)
        rho = torch.empty((1, D, D), dtype=torch.complex128, device=dev)

        def step():
            eng.partial_trace_sv_t(psi, traced, out=rho)

        step()
        torch.cuda.synchronize()
        ms, _, _, launches = self.time_steps(step, steps, 2, profile=False)
        flops = 8.0 * D * D * T  # algorithmic (no Hermitian discount)
        blocks = D // 32
        executed = 8.0 * T * 32 * 32 * (blocks * (blocks + 1) // 2)  # the kernel computes the blocks on and above the diagonal
        bytes_ = 16.0 * (D * T + D * D)
        # parity: 64 random elements of rho against a direct fp64 dot product (torch, same device)
        Psi = psi.view(T, D)  # index = t * D + i  (kept bits are the low ones)
        ii = torch.randint(0, D, (64,), generator=torch.Generator().manual_seed(1))
        jj = torch.randint(0, D, (64,), generator=torch.Generator().manual_seed(2))
        ref = torch.stack([(Psi[:, i] * Psi[:, j].conj()).sum() for i, j in zip(ii.tolist(), jj.tolist())])
        got = rho[0][ii.to(dev), jj.to(dev)]
        err = float((got - ref).abs().max().item() / rho.abs().max().item())
        peaks = load_peaks()
        out = {"workload": f"k1: partial trace of one statevector of {n_qubits} qubits to {n_keep} kept (Psi {D} x {T}, the synthetic "
                           "large case of SURVEY 8d: two cutoff-64 modes kept)", "ms_per_step": ms, "gpu_launches_per_step": launches,
               "kernel": "trace_sv_dmma_kernel<4> (mma.sync m8n8k4 f64, 4 x 4 tiles per warp, blocks on and above the diagonal + mirrored stores)",
               "flops": flops, "tflops": flops / (ms * 1e-3) / 1e12, "algorithmic_frac_of_fp64_peak": flops / (ms * 1e-3) / 1e12 / self.fp64_peak,
               "executed_flops": executed, "executed_tflops": executed / (ms * 1e-3) / 1e12,
               "frac_of_fp64_peak": executed / (ms * 1e-3) / 1e12 / self.fp64_peak,
               "bytes": bytes_, "gbs": bytes_ / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": bytes_ / (ms * 1e-3) / 1e9 / peaks[0],
               "parity_max_abs_drho_over_max_rho_64_elements": err}
        del psi, rho
        torch.cuda.empty_cache()
        return out


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


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

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # NCCL prints its version banner on stdout; keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    b = Bench(args, rank, world, local)
    wl = make_workload(args.config, args.frames)
    cfg = describe(wl, world, args)
    scaling = cfg.pop("_scaling")
    mode = sharding_mode(wl, world, args)
    heavy = wl["N"] >= 512  # a point-by-point step of c5 takes ~2.8 s: two of them
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    main_res = b.frames_workload(wl, args.steps, args.warmup, mode, pbp_steps=0 if args.no_point_by_point else (2 if heavy else args.steps),
                                 want_cpu=not args.no_cpu_baseline, cpu_seconds=args.cpu_seconds,
                                 e2e_steps=min(args.steps, 10) if heavy else args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- legs: the other named configurations, timed the same way ----
    legs = {}
    for name in [s for s in args.legs.split(",") if s]:
        if name == args.config:
            continue
        try:
            if name == "k1":
                if world == 1:
                    legs["k1"] = b.k1_leg()
                continue
            lw = make_workload(name, 512 if name == "c3" else None)
            lmode = sharding_mode(lw, world, args)
            if lmode == "grid":
                lmode = "replicas"  # small single-frame legs: every GPU its own copy (weak)
            big = name == "c3"
            leg_sampler = ClockSampler(local) if rank == 0 else None  # (the legs run on a GPU the main workload has heated up)
            if leg_sampler:
                leg_sampler.start()
            try:
                r = b.frames_workload(lw, 10 if big else 20, 6 if big else 10, lmode, pbp_steps=0, want_cpu=False,
                                      e2e_steps=3 if big else 20)
            finally:
                leg_clocks = leg_sampler.stop() if leg_sampler else None
            r["clocks"] = leg_clocks
            r["workload"] = WORKLOAD_TEXT[name] + (f", 512 frames over {world} GPU(s)" if name == "c3" else "")
            r["scaling"] = "weak (replicas)" if lmode == "replicas" else ("strong" if world > 1 else "single GPU")
            legs[name] = r
        except Exception as exc:  # a leg must not take the headline line down with it
            legs[name] = {"error": repr(exc)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, hbm_src = load_peaks()
    roof = main_res["roofline"]
    if roof:
        roof["hbm_store"].update({"peak": hbm_peak, "peak_source": hbm_src, "frac": roof["hbm_store"]["achieved"] / hbm_peak})
        traffic = None  # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                t = json.load(f).get(wl["name"])
            if t:
                traffic = {"bytes": t["dram_bytes_read"] + t["dram_bytes_write"], "source": t["source"]}
        except Exception:
            pass
        roof["traffic"] = traffic
    out = {
        "metric": METRIC, "value": main_res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": cfg, "e2e": main_res["e2e"],
        "gpu_launches": int(round(main_res["gpu_launches_per_step"] * args.steps)), "clocks": clocks,
        "roofline": roof, "cpu_baseline": main_res.get("cpu_baseline"), "parity": main_res.get("parity"),
        "gather": None if world == 1 else {"grid": "torch.distributed.reduce (NCCL, sum of disjoint shares) to rank 0 inside the step",
                                             "frames": "none (results stay on the GPU that made them); e2e assembles ONE host array",
                                             "replicas": "none"}[mode],
        "legs": legs,
    }
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
