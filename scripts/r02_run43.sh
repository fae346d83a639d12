#!/bin/bash
# round 2, GPU run 43: the bench line at N = 1 on the final build
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r02f_bench_n1.err
