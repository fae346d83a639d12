#!/bin/bash
# round 2, GPU run 6: REDUX prefixes, device-side -stats, matchIndexed cross-check
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_match_indexed_cross_check.py tests/test_gpu_cli.py tests/test_gpu_zz_shim_extras.py -x -q > gpurun_out/r02_parity6.log 2>&1
echo "parity rc=$?"; tail -6 gpurun_out/r02_parity6.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-configs --no-cli > gpurun_out/r02_bench6.json 2> gpurun_out/r02_bench6.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench6.err; python -c "
import json; b=json.loads(open('gpurun_out/r02_bench6.json').read().strip().splitlines()[-1]); print(b.get('ms_per_step'), b.get('kernel_ms'), b.get('roofline',{}).get('frac'), b.get('e2e',{}).get('ms_per_step'), b.get('parity_gate',{}).get('ok'))"
