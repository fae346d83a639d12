#!/bin/bash
# round 2, GPU run 32: combine scan with block prefix / suffix maxima: parity of the window-parallel paths, timings
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "window_parallel or combine or shard or golden" > gpurun_out/r02_parity32.log 2>&1
echo "parity rc=$?"; tail -4 gpurun_out/r02_parity32.log
timeout 300 python scripts/time_variants.py > gpurun_out/r02_combine_variants.log 2>&1; cat gpurun_out/r02_combine_variants.log
