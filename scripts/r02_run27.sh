#!/bin/bash
# round 2, GPU run 27: head candidates whose upward extension is known from the divergence update: parity, timings
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "not wide and not global and not query" > gpurun_out/r02_parity27.log 2>&1
echo "parity rc=$?"; tail -4 gpurun_out/r02_parity27.log
timeout 300 python scripts/time_variants.py > gpurun_out/r02_headext_variants.log 2>&1; cat gpurun_out/r02_headext_variants.log
