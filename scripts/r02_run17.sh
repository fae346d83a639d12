#!/bin/bash
# round 2, GPU run 17: a[] staged in shared memory between the phases of the wide gather build (A/B), wide parity
mkdir -p gpurun_out
for lib in pbwt_b200/libpbwtgpu_prev.so pbwt_b200/libpbwtgpu.so; do
  echo "== $lib" >> gpurun_out/r02_ab_stage_a.log
  PBWTGPU_LIB=$PWD/$lib timeout 300 python scripts/time_wide3.py >> gpurun_out/r02_ab_stage_a.log 2>&1
done
cat gpurun_out/r02_ab_stage_a.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "wide or global or shard" > gpurun_out/r02_parity17.log 2>&1
echo "parity rc=$?"; tail -4 gpurun_out/r02_parity17.log
