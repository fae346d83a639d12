"""Development aid (torchrun, N ranks): host wall time of the phases of one sharded build + maxWithin step of config 2's
shape -- prepare (pass 1 + combine + aggregate), export (+ stream sync), all-gather, finish (prefix, pass 2, outputs)."""
import os, sys, time
sys.path.insert(0, '.')
import torch, torch.distributed as dist
import pbwt_b200 as P
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
M, N = 10000, 100000
eng = P.Engine(local); eng.profile(True)
stream = torch.cuda.Stream()
eng.set_stream(stream.cuda_stream)
flags = P.WANT_YZ | P.WANT_MAXWITHIN
with torch.cuda.stream(stream):
    panel = eng.panel_synthetic(M, N, seed=0x5EED0002, L=2000, mu=2.0, first_site=rank * N)
    ai = eng.shard_aggregate_ints(M)
    agg = torch.zeros(ai, dtype=torch.int32, device="cuda"); aggs = torch.zeros((world, ai), dtype=torch.int32, device="cuda")
    acc = [0.0] * 6
    reps = 12
    for rep in range(reps + 3):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        sh = eng.shard_prepare(panel, rank * N, world * N, flags); t1 = time.perf_counter()
        torch.cuda.synchronize(); t1s = time.perf_counter()
        sh.export(agg.data_ptr()); t2 = time.perf_counter()
        dist.all_gather_into_tensor(aggs, agg); t3 = time.perf_counter()
        torch.cuda.synchronize(); t3s = time.perf_counter()
        res = sh.finish(aggs.data_ptr(), rank, flags); t4 = time.perf_counter()
        torch.cuda.synchronize(); t5 = time.perf_counter()
        sh.free(); res.free()
        tm = eng.last_timings()
        if rep >= 3:
            for j, v in enumerate([t1 - t0, t1s - t1, t2 - t1s, t3s - t2, t4 - t3s, t5 - t4]): acc[j] += v
    names = ['prepare call', 'prepare drain', 'export', 'all-gather (+sync)', 'finish call', 'finish drain']
    print('rank', rank, {n: round(v / reps * 1e3, 3) for n, v in zip(names, acc)}, 'sum', round(sum(acc) / reps * 1e3, 3), 'kernels', {k: (round(v[0], 3), v[1]) for k, v in tm.items() if v[1]}, flush=True)
eng.set_option("window_times", 1)
with torch.cuda.stream(stream):
    sh = eng.shard_prepare(panel, rank * N, world * N, flags); sh.export(agg.data_ptr()); dist.all_gather_into_tensor(aggs, agg)
    res = sh.finish(aggs.data_ptr(), rank, flags); torch.cuda.synchronize(); sh.free(); res.free()
dist.destroy_process_group()
