#!/bin/bash
# round 2, GPU run 7: pack3 fused into the sweep (dev build): parity, then A/B against the current build
mkdir -p gpurun_out
PBWTGPU_LIB=$PWD/pbwt_b200/libpbwtgpu_dev.so timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r02_parity7.log 2>&1
echo "parity(dev) rc=$?"; tail -6 gpurun_out/r02_parity7.log
for lib in pbwt_b200/libpbwtgpu.so pbwt_b200/libpbwtgpu_dev.so pbwt_b200/libpbwtgpu.so pbwt_b200/libpbwtgpu_dev.so; do
  PBWTGPU_LIB=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-configs --no-cli --no-cpu-baseline > gpurun_out/r02_bench7.json 2> gpurun_out/r02_bench7.err
  echo "$lib rc=$?"; python -c "
import json; b=json.loads(open('gpurun_out/r02_bench7.json').read().strip().splitlines()[-1]); print(b.get('ms_per_step'), b.get('kernel_ms'), b.get('e2e',{}).get('ms_per_step'), b['e2e']['ms_per_step_one_in_flight'])"
done
