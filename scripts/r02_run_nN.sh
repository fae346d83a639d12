#!/bin/bash
# round 2: the bench line at N GPUs (weak config-2 shape + parity, configs 4 and 5 site-sharded); usage: r02_run_nN.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02f_bench_n$N.json 2> gpurun_out/r02f_bench_n$N.err
echo "bench n$N rc=$?"; tail -c 800 gpurun_out/r02f_bench_n$N.err | grep -v Warning | tail -5; tail -c 400 gpurun_out/r02f_bench_n$N.json
