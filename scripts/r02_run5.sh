#!/bin/bash
# round 2, GPU run 5 (2 GPUs): the bench line at N = 2 (weak C2 + parity, config 4 and config 5 site-sharded)
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
echo "bench n2 rc=$?"; tail -c 1500 gpurun_out/r02_bench_n2.err; tail -c 3000 gpurun_out/r02_bench_n2.json
