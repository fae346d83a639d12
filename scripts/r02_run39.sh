#!/bin/bash
# round 2, GPU run 39 (2 GPUs): where a sharded step's time goes on the host side, per-window times of pass 2 per rank
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 scripts/shard_phases.py > gpurun_out/r02_shard_phases.log 2>&1
grep "^rank" gpurun_out/r02_shard_phases.log || tail -20 gpurun_out/r02_shard_phases.log
grep "window times" gpurun_out/r02_shard_phases.log | python -c "
import re,sys,statistics
for line in sys.stdin:
    rows=[(int(a),int(b),int(c)) for a,b,c in re.findall(r' (\d+):(\d+)/(\d+)', line)]
    t=[r[1] for r in rows]
    print(len(rows),'windows first',rows[0],'second',rows[1],'last',rows[-1],'mean',round(statistics.mean(t)),'max',max(t),'top',sorted(rows,key=lambda r:-r[1])[:5])
"
