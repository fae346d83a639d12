#!/bin/bash
# round 2, GPU run 34: the whole GPU suite on the current build, the bench line, launch list and ncu --set full captures
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r02_final_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r02f_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_bench_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-configs --no-cli --no-cpu-baseline > gpurun_out/r02_ncu_launch.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:.*sweep_smem_kernel.*' -s 0 -c 1 -o gpurun_out/r02f_c2_pass2 python scripts/prof_run.py 10000 100000 5 > gpurun_out/r02_ncu_pass2.log 2>&1
echo "ncu pass2 rc=$?"; tail -2 gpurun_out/r02_ncu_pass2.log
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:combine_sort_kernel' -s 2 -c 1 -o gpurun_out/r02f_c2_combine python scripts/prof_run.py 10000 100000 5 > gpurun_out/r02_ncu_combine.log 2>&1
echo "ncu combine rc=$?"; tail -2 gpurun_out/r02_ncu_combine.log
