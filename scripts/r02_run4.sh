#!/bin/bash
# round 2, GPU run 4: batch sizes of the wide kernel (A/B against the previous build), wide parity, C5 timing
mkdir -p gpurun_out
for lib in pbwt_b200/libpbwtgpu_prev.so pbwt_b200/libpbwtgpu.so; do
  echo "== $lib" >> gpurun_out/r02_ab_batch.log
  PBWTGPU_LIB=$PWD/$lib timeout 300 python scripts/time_wide3.py >> gpurun_out/r02_ab_batch.log 2>&1
done
cat gpurun_out/r02_ab_batch.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "wide or global or query or shard" > gpurun_out/r02_parity4.log 2>&1
echo "parity rc=$?"; tail -4 gpurun_out/r02_parity4.log
timeout 600 python scripts/time_c5.py 100000 0 > gpurun_out/r02_time_c5c.log 2>&1; tail -2 gpurun_out/r02_time_c5c.log
