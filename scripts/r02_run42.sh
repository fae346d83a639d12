#!/bin/bash
# round 2, GPU run 42: edge-window lengths chosen by load: parity, bench line at N = 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -x -q > gpurun_out/r02_parity42.log 2>&1
echo "parity rc=$?"; tail -3 gpurun_out/r02_parity42.log
timeout 300 python scripts/time_variants.py 2>&1 | tail -2
timeout 300 python scripts/window_times.py 10000 100000 2>&1 | tail -2
