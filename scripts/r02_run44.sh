#!/bin/bash
# round 2, GPU run 44: compute-sanitizer racecheck + memcheck on the small round-1 and round-2 targets
mkdir -p gpurun_out
rm -f gpurun_out/r02_sanitizer*.log
for t in race_small race_small2; do
  for tool in racecheck memcheck; do
    echo "== $tool $t" >> gpurun_out/r02_sanitizer.log
    timeout 900 compute-sanitizer --tool $tool python scripts/$t.py > gpurun_out/r02_sanitizer_${tool}_$t.log 2>&1
    grep "Race reported\|and Read\|and Write\|SUMMARY\|^ok\|Error\|Invalid" gpurun_out/r02_sanitizer_${tool}_$t.log | sed -E 's/void pbwtgpu::<unnamed>:://g' | cut -c1-260 | sort | uniq -c | sort -rn | head -30 >> gpurun_out/r02_sanitizer.log
  done
done
cat gpurun_out/r02_sanitizer.log
