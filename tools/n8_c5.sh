TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655"
timeout 200 $TR tests/multi_gpu_check.py > gpurun_out/multi_check_n8.log 2>&1; echo check rc=$?
timeout 200 $TR bench.py --gpus 8 --config c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1d_bench_n8_c5.json 2> gpurun_out/n8_c5.err; echo c5 rc=$?
