# round-2 evidence (run under gpurun, one GPU): launch list of the default bench command, ncu --set full captures of the
# two dominant kernels (c5: parked ring kernel; c2 leg: resident 4-unit kernel) with source, raw metrics to text
set -x
export BQ_B200_SASS_VERIFY=0   # the load-time self-check launches every kernel variant twice: keep it out of the captures
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-point-by-point --fp64-probe-seconds 0.02"
$B --legs c1,c2,c3,c4 > gpurun_out/r2_plain_run.json 2> gpurun_out/r2_plain_run.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c5.csv $B --legs c1,c2,c3,c4 > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wigner_sym_diag_parked -s 4 -c 1 -f -o gpurun_out/r2_sym_parked_c5 $B --legs "" > gpurun_out/ncu_f5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wigner_sym_resident -s 4 -c 1 -f -o gpurun_out/r2_sym_c2 $B --config c2 --legs "" > gpurun_out/ncu_f2.log 2>&1
ls -la gpurun_out/*.ncu-rep
