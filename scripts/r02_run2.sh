#!/bin/bash
# round 2, GPU run 2: window-parallel query sweep (parity + timing), ncu of the wide-panel pass 2, scan-by-distance A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "query" > gpurun_out/r02_query_tests.log 2>&1
echo "query tests rc=$?"; tail -15 gpurun_out/r02_query_tests.log
timeout 600 python scripts/time_c5.py 100000 0 > gpurun_out/r02_time_c5.log 2>&1
echo "time_c5 rc=$?"; cat gpurun_out/r02_time_c5.log | tail -5
timeout 900 python -m pytest tests/test_gpu_full_size.py -x -q -s -k config5 > gpurun_out/r02_fullsize_c5.log 2>&1
echo "c5 full rc=$?"; tail -5 gpurun_out/r02_fullsize_c5.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
   -k 'regex:.*sweep_global_kernel<.int.0, .bool.0, .int.0, .bool.0.*' -s 1 -c 1 -o gpurun_out/r02_c3_pass2 python scripts/prof_c3.py > gpurun_out/r02_ncu_c3.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02_ncu_c3.log
for lib in pbwt_b200/libpbwtgpu.so pbwt_b200/libpbwtgpu_scandist.so pbwt_b200/libpbwtgpu.so pbwt_b200/libpbwtgpu_scandist.so; do
  timeout 200 python scripts/ab_lib.py $lib >> gpurun_out/r02_ab_scandist.log 2>&1
done
cat gpurun_out/r02_ab_scandist.log
timeout 300 python scripts/window_times.py 10000 100000 > gpurun_out/r02_window_times.log 2>&1; cat gpurun_out/r02_window_times.log
