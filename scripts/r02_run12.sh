#!/bin/bash
# round 2, GPU run 12: lane-parallel fused pack3 + immediate-offset addressing in the wide kernel (dev build): parity, A/B
mkdir -p gpurun_out
PBWTGPU_LIB=$PWD/pbwt_b200/libpbwtgpu_dev2.so timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r02_parity12.log 2>&1
echo "parity(dev2) rc=$?"; tail -6 gpurun_out/r02_parity12.log
for lib in pbwt_b200/libpbwtgpu.so pbwt_b200/libpbwtgpu_dev2.so pbwt_b200/libpbwtgpu.so pbwt_b200/libpbwtgpu_dev2.so; do
  PBWTGPU_LIB=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-configs --no-cli --no-cpu-baseline > gpurun_out/r02_bench12.json 2> gpurun_out/r02_bench12.err
  echo "$lib rc=$?"; python -c "
import json; b=json.loads(open('gpurun_out/r02_bench12.json').read().strip().splitlines()[-1]); print(b.get('ms_per_step'), b.get('kernel_ms'), b.get('e2e',{}).get('ms_per_step'), b['e2e']['ms_per_step_one_in_flight'])"
done
