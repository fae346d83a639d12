#!/bin/bash
# round 2, GPU run 55: the whole GPU suite on the final build
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r02_final_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
