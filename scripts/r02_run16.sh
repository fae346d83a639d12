#!/bin/bash
# round 2, GPU run 16: grouped records (parity, e2e), CLI tests through the shim's grouped path
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py tests/test_gpu_zz_shim_extras.py -x -q -k "grouped or cli or shim or stats or golden" > gpurun_out/r02_parity16.log 2>&1
echo "parity rc=$?"; tail -5 gpurun_out/r02_parity16.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-configs > gpurun_out/r02_bench16.json 2> gpurun_out/r02_bench16.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench16.err; python -c "
import json; b=json.loads(open('gpurun_out/r02_bench16.json').read().strip().splitlines()[-1]); print(b.get('ms_per_step'), b.get('e2e'), b.get('parity_gate'), b.get('cli'))"
