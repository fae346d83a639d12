"""torchrun --nproc-per-node R scripts/multi_gpu_check.py : the NCCL site-sharded build must equal the single-GPU
build of the same panel (which the parity suite pins to the reference)."""
import os, sys
sys.path.insert(0, '.')
import numpy as np, torch, torch.distributed as dist
import pbwt_b200 as P

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
torch.cuda.set_device(lr)
eng = P.Engine(lr)
stream = torch.cuda.Stream()  # explicit stream: kernels and the NCCL all-gather are ordered on it (handle 0 would mean "own stream")
torch.cuda.set_stream(stream)
eng.set_stream(stream.cuda_stream)
M, Ns = 6000, 4000
flags = P.WANT_YZ | P.WANT_MAXWITHIN
pnl = eng.panel_synthetic(M, Ns, seed=0x5EED0004, L=500, mu=3.0, first_site=rank * Ns)
sh = eng.shard_prepare(pnl, rank * Ns, Ns * world, flags)
ai = eng.shard_aggregate_ints(M)
agg = torch.zeros(ai, dtype=torch.int32, device="cuda"); aggs = torch.zeros((world, ai), dtype=torch.int32, device="cuda")
sh.export(agg.data_ptr()); dist.all_gather_into_tensor(aggs, agg)
res = sh.finish(aggs.data_ptr(), rank, flags); out = res.fetch(); res.free(); sh.free()
outs = [None] * world
dist.all_gather_object(outs, (out["yz"], out["a_end"], out["d_end"], out["matches"]))
if rank == 0:
    full = eng.panel_synthetic(M, Ns * world, seed=0x5EED0004, L=500, mu=3.0)
    r = eng.build_device(full, flags); ref = r.fetch(); r.free()
    ok = (np.concatenate([o[0] for o in outs]).tobytes() == ref["yz"].tobytes() and (outs[-1][1] == ref["a_end"]).all()
          and (outs[-1][2] == ref["d_end"]).all() and (np.concatenate([o[3] for o in outs]) == ref["matches"]).all())
    print("MULTI_GPU_CHECK", "OK" if ok else "MISMATCH", "world", world, "records", len(ref["matches"]), "yz", ref["yz"].size, flush=True)
dist.destroy_process_group()
