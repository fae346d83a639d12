#!/bin/bash
# round 2, GPU run 29: edge-window sizes against pass 2, per-window times
mkdir -p gpurun_out
timeout 300 python scripts/time_edges.py > gpurun_out/r02_edges.log 2>&1; cat gpurun_out/r02_edges.log
timeout 300 python scripts/window_times.py 10000 100000 2>&1 | tail -3
