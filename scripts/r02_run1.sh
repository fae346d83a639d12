#!/bin/bash
# round 2, GPU run 1: full-size parity against the compiled reference, the new bench line, ncu of the wide-panel pass 2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02_run1_smi.txt 2>&1
lscpu | head -20 >> gpurun_out/r02_run1_smi.txt; free -g >> gpurun_out/r02_run1_smi.txt
timeout 1500 python -m pytest tests/test_gpu_full_size.py -x -q -s > gpurun_out/r02_fullsize.log 2>&1
echo "fullsize rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r02_bench1.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02_bench1_ref.json 2> gpurun_out/r02_bench1_ref.err
echo "ref rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
   -k 'regex:.*sweep_global_kernel<0, 0, 0, 0.*' -s 1 -c 1 -o gpurun_out/r02_c3_pass2 python scripts/prof_c3.py > gpurun_out/r02_ncu_c3.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02_ncu_c3.log
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_full_size.py > gpurun_out/r02_pytest_gpu1.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu1.log
tail -5 gpurun_out/r02_fullsize.log
