# round-1 (r1d) evidence: launch list + ncu --set full captures of the dominant kernels (run under gpurun, one GPU)
set -x
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --fp64-probe-seconds 0.02"
$B > gpurun_out/r1d_plain_run.json 2> gpurun_out/r1d_plain_run.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1d_launches_c2.csv $B > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wigner_sym_resident -s 3 -c 1 -f -o gpurun_out/r1d_sym_c2 $B > gpurun_out/ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wigner_sym_diag -s 1 -c 1 -f -o gpurun_out/r1d_sym_c5 python bench.py --config c5 --steps 1 --warmup 3 --no-cpu-baseline --fp64-probe-seconds 0.02 > gpurun_out/ncu_f5.log 2>&1
python - <<'PY' > gpurun_out/r1d_fft_timing.txt 2>&1
import time, json, numpy as np, torch, sys
sys.path.insert(0, '.')
import workloads
from bosonic_qiskit_b200 import get_engine
eng = get_engine(0)
dev = torch.device('cuda', 0)
for N, S in ((16, 200), (64, 1024), (64, 2048), (256, 2048)):
    rho = torch.from_numpy(workloads.random_rho(N, seed=N, rank=4)).to(dev)
    x = np.linspace(-6, 6, S)
    for _ in range(2):
        eng.wigner_fft_t(rho, x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        eng.wigner_fft_t(rho, x)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(json.dumps({"method": "fft", "cutoff": N, "grid": S, "ms": round(ms, 4), "points_per_s": S * S / ms * 1e3,
                      "dft_dfma_tflops": 2.0 * S ** 3 / (ms * 1e-3) / 1e12}))
PY
