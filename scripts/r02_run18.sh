#!/bin/bash
# round 2, GPU run 18: a-staging restricted to a-only builds (timing), wide + CLI parity, bench line, ncu --set full of config 2's pass 2
mkdir -p gpurun_out
timeout 300 python scripts/time_wide3.py > gpurun_out/r02_stage_a_final.log 2>&1; cat gpurun_out/r02_stage_a_final.log
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -x -q -k "wide or global or shard or cli or CLI or match" > gpurun_out/r02_parity18.log 2>&1
echo "parity rc=$?"; tail -4 gpurun_out/r02_parity18.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-configs > gpurun_out/r02d_bench_n1_noconfigs.json 2> gpurun_out/r02d_bench_n1_noconfigs.err
echo "bench rc=$?"; cat gpurun_out/r02d_bench_n1_noconfigs.json | head -c 3000
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:.*sweep_smem_kernel.*' -s 0 -c 2 -o gpurun_out/r02_c2_pass2 python scripts/prof_run.py 10000 100000 5 > gpurun_out/r02_ncu_pass2.log 2>&1
echo "ncu pass2 rc=$?"; tail -3 gpurun_out/r02_ncu_pass2.log
