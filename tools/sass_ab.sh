export PYTHONPATH=$PWD
run() { timeout 200 python bench.py --config $1 --steps 20 --warmup 6 --no-cpu-baseline --no-point-by-point --legs "" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$2', '$1', round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), round(d['roofline']['frac'],4))"; }
for rep in 1 2; do
run c3 default
BQ_B200_SASS=plain run c3 plain
BQ_B200_LIBRARY=$PWD/bosonic_qiskit_b200/libbq_b200_blocks.so run c3 tails
done
run c2 default
BQ_B200_SASS=plain run c2 plain
BQ_B200_LIBRARY=$PWD/bosonic_qiskit_b200/libbq_b200_blocks.so run c2 tails
BQ_B200_SASS=plain run c5 plain
