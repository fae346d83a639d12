#!/bin/bash
# round 2, GPU run 26: after the pass-2 experiments were reverted: quick parity, bench line with the CLI phase times
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -x -q -k "not wide and not global and not query" > gpurun_out/r02_parity26.log 2>&1
echo "parity rc=$?"; tail -4 gpurun_out/r02_parity26.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-configs > gpurun_out/r02e_bench_n1_noconfigs.json 2> gpurun_out/r02e_bench_n1_noconfigs.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02e_bench_n1_noconfigs.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], json.dumps(d.get('cli')), d['roofline']['kernel_ms_per_step_by_role'])
PY
