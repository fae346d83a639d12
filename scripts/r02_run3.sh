#!/bin/bash
# round 2, GPU run 3: combine as a sort, query fast path, scan by distance: parity + timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r02_parity3.log 2>&1
echo "parity rc=$?"; tail -8 gpurun_out/r02_parity3.log
timeout 600 python scripts/time_c5.py 100000 0 > gpurun_out/r02_time_c5b.log 2>&1
echo "time_c5 rc=$?"; tail -3 gpurun_out/r02_time_c5b.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-configs --no-cli > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench3.err; python -c "
import json; b=json.loads(open('gpurun_out/r02_bench3.json').read().strip().splitlines()[-1]); print(b.get('ms_per_step'), b.get('kernel_ms'), b.get('roofline',{}).get('frac'), b.get('e2e',{}).get('ms_per_step'), b.get('parity_gate',{}).get('ok'))"
timeout 900 python -m pytest tests/test_gpu_full_size.py -x -q -s -k "config2 or config5" > gpurun_out/r02_fullsize3.log 2>&1
echo "full rc=$?"; tail -6 gpurun_out/r02_fullsize3.log
