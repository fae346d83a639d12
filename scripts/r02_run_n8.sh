#!/bin/bash
# round 2: the bench line at N = 8 (weak config-2 shape + parity, configs 4 and 5 site-sharded)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
echo "bench n8 rc=$?"; tail -c 800 gpurun_out/r02_bench_n8.err | grep -v Warning | tail -5; tail -c 400 gpurun_out/r02_bench_n8.json
