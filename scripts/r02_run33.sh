#!/bin/bash
# round 2, GPU run 33: A/B of the radix sorts' in-warp ranking: match.any vs eight votes (pass 1, combine; wide sorts)
mkdir -p gpurun_out
for lib in pbwt_b200/libpbwtgpu.so pbwt_b200/libpbwtgpu_dev.so pbwt_b200/libpbwtgpu.so pbwt_b200/libpbwtgpu_dev.so; do
  echo "== $lib" >> gpurun_out/r02_ab_match_any.log
  PBWTGPU_LIB=$PWD/$lib timeout 300 python scripts/time_variants.py 2>&1 | tail -1 >> gpurun_out/r02_ab_match_any.log
done
for lib in pbwt_b200/libpbwtgpu.so pbwt_b200/libpbwtgpu_dev.so; do
  echo "== $lib" >> gpurun_out/r02_ab_match_any.log
  PBWTGPU_LIB=$PWD/$lib timeout 300 python scripts/time_wide3.py >> gpurun_out/r02_ab_match_any.log 2>&1
done
cat gpurun_out/r02_ab_match_any.log
PBWTGPU_LIB=$PWD/pbwt_b200/libpbwtgpu_dev.so timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "window_parallel or combine or shard or golden or wide or global" > gpurun_out/r02_parity33.log 2>&1
echo "parity(dev) rc=$?"; tail -3 gpurun_out/r02_parity33.log
