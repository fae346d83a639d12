#!/bin/bash
# round 2, GPU run 35: which piece of the match bookkeeping costs what (development builds, results of the scans are wrong in them)
mkdir -p gpurun_out
for v in "" _NO_WMAX _NO_DETECT _NO_ROUNDS _ALL3; do
  echo "== libpbwtgpu$v.so" >> gpurun_out/r02_bookkeeping.log
  PBWTGPU_LIB=$PWD/pbwt_b200/libpbwtgpu$v.so timeout 200 python scripts/time_nocand.py 2>&1 | tail -1 >> gpurun_out/r02_bookkeeping.log
done
cat gpurun_out/r02_bookkeeping.log
