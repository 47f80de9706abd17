# scaling evidence for N GPUs (run under `gpurun --gpus N`): default config (weak) and config 5 as one grid (strong)
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655"
timeout 200 $TR bench.py --gpus $N --no-cpu-baseline > gpurun_out/r1d_scale_n${N}_c2.json 2> gpurun_out/scale_n${N}_c2.err; echo c2 rc=$?
timeout 200 $TR bench.py --gpus $N --config c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1d_scale_n${N}_c5.json 2> gpurun_out/scale_n${N}_c5.err; echo c5 rc=$?
