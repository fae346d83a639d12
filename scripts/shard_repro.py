"""Development aid: the multi-GPU protocol of scripts/multi_gpu_check.py run shard by shard on one GPU."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
import pbwt_b200 as P
eng = P.Engine(0)
M, Ns, world = 6000, 4000, int(sys.argv[1]) if len(sys.argv) > 1 else 8
flags = P.WANT_YZ | P.WANT_MAXWITHIN
ai = eng.shard_aggregate_ints(M)
aggs = torch.zeros((world, ai), dtype=torch.int32, device="cuda")
panels = [eng.panel_synthetic(M, Ns, seed=0x5EED0004, L=500, mu=3.0, first_site=r * Ns) for r in range(world)]
shards = [eng.shard_prepare(panels[r], r * Ns, Ns * world, flags) for r in range(world)]
for r in range(world):
    shards[r].export(aggs[r].data_ptr())
outs = []
for r in range(world):
    res = shards[r].finish(aggs.data_ptr(), r, flags); outs.append(res.fetch()); res.free()
full = eng.panel_synthetic(M, Ns * world, seed=0x5EED0004, L=500, mu=3.0)
rr = eng.build_device(full, flags); ref = rr.fetch(); rr.free()
yz = np.concatenate([o["yz"] for o in outs]); rec = np.concatenate([o["matches"] for o in outs])
print("yz ok", yz.tobytes() == ref["yz"].tobytes(), "a ok", (outs[-1]["a_end"] == ref["a_end"]).all(), "d ok", (outs[-1]["d_end"] == ref["d_end"]).all(),
      "rec ok", len(rec) == len(ref["matches"]) and (rec == ref["matches"]).all(), len(rec), len(ref["matches"]))
if yz.tobytes() != ref["yz"].tobytes():
    # first differing shard
    off = 0
    for r, o in enumerate(outs):
        n = len(o["yz"]); same = o["yz"].tobytes() == ref["yz"][off:off + n].tobytes(); print("shard", r, "yz same", same, n); off += n
