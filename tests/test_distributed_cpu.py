"""CPU tests of the multi-GPU host logic with world_size 2 and 3 over gloo.

The device steps are stood in for by the CPU oracle (tests may use it as the checker/stand-in); what is under
test is the distributed algorithm the GPU ranks run (include/pbwtgpu.h, multi-GPU section): sites are
partitioned across ranks, every rank reduces its shard to an aggregate node (order + divergence + tie groups),
ONE all-gather exchanges the aggregates, each rank combines the earlier ones into its exact start state, and the
concatenation of the per-rank outputs equals the sequential sweep."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O


def aggregate_node(x, M, N, k0, a_start):
    """A shard swept on its own from the panel's start order: order, divergence (0 = never separated in this
    shard, else the global label), tie groups and group-head divergence (windows.cu: make_groups)."""
    _, a_end, d_end, _ = O.sweep_from(x, M, N, k0, k0 + N + 1, a_start, None)  # ntotal beyond: no final report
    d = d_end.copy()
    d[0] = d[M] = 0
    flag = np.ones(M, bool)
    flag[1:] = d[1:M] != 0
    rank = np.cumsum(flag) - 1
    g = np.zeros(M, np.int32)
    g[a_end] = rank
    gfd = np.zeros(M, np.int32)
    gfd[rank[flag]] = d[:M][flag]
    return np.concatenate([a_end, d, g, gfd]).astype(np.int32)  # ord[M] | dv[M+1] | g[M] | gfd[M]


def split(agg, M):
    return agg[:M], agg[M:2 * M + 1], agg[2 * M + 1:3 * M + 1], agg[3 * M + 1:]


def combine(I, J, M):
    """I (+) J (windows.cu): order(I) stably re-sorted by J's group; divergence = J's group-head divergence
    between J-groups, else the max of I's divergence over the I-positions in between."""
    ordI, dvI, _, _ = split(I, M)
    _, _, gJ, gfdJ = split(J, M)
    key = gJ[ordI]
    perm = np.argsort(key, kind="stable")  # positions of order(I), re-sorted
    ordN = ordI[perm]
    dvN = np.zeros(M + 1, np.int32)
    for i in range(1, M):
        u, v = ordN[i - 1], ordN[i]
        if gJ[u] != gJ[v]:
            dvN[i] = gfdJ[gJ[v]]
        else:
            lo, hi = perm[i - 1], perm[i]
            dvN[i] = dvI[lo + 1:hi + 1].max()
    flag = np.ones(M, bool)
    flag[1:] = dvN[1:M] != 0
    rank = np.cumsum(flag) - 1
    g = np.zeros(M, np.int32)
    g[ordN] = rank
    gfd = np.zeros(M, np.int32)
    gfd[rank[flag]] = dvN[:M][flag]
    return np.concatenate([ordN, dvN, g, gfd]).astype(np.int32)


def worker(rank, world, port, M, N, seed, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(seed)
    hap = (rng.random((M, N)) < 0.2).astype(np.uint8)
    hap[: M // 3] = hap[0]
    a_start = rng.permutation(M).astype(np.int32)
    x = O.pack_panel(hap)
    bounds = [N * r // world for r in range(world + 1)]
    k0, k1 = bounds[rank], bounds[rank + 1]
    # 1. local aggregate; 2. the one exchange: all-gather of aggregates in site order
    agg = torch.from_numpy(aggregate_node(x[k0:k1], M, k1 - k0, k0, a_start))
    gathered = [torch.zeros_like(agg) for _ in range(world)]
    dist.all_gather(gathered, agg)
    # 3. exact start state of this rank = combination of the earlier shards' aggregates
    a0, d0 = a_start, None
    if rank > 0:
        P = gathered[0].numpy()
        for r in range(1, rank):
            P = combine(P, gathered[r].numpy(), M)
        a0, d0 = split(P, M)[0], split(P, M)[1]
    yz, a_end, d_end, rec = O.sweep_from(x[k0:k1], M, k1 - k0, k0, N, a0, d0)
    outs = [None] * world
    dist.all_gather_object(outs, (yz, a_end, d_end, rec))
    if rank == 0:
        orc = O.port()
        yz_ref, a_ref, d_ref = orc.build(x, M, N, a_start, with_d=True)
        rec_ref, n = orc.max_within(yz_ref, M, N, a_start)
        ok = (np.concatenate([o[0] for o in outs]).tobytes() == yz_ref.tobytes()
              and (outs[-1][1] == a_ref).all() and (outs[-1][2] == d_ref).all()
              and len(np.concatenate([o[3] for o in outs])) == n and (np.concatenate([o[3] for o in outs]) == rec_ref).all())
        q.put(bool(ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,M,N", [(2, 40, 60), (3, 64, 45), (2, 7, 9), (4, 33, 30)])
def test_site_sharded_build_gloo(world, M, N):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000) + world
    procs = [ctx.Process(target=worker, args=(r, world, port, M, N, 1234 + world, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


@pytest.mark.parametrize("M,N,seed", [(40, 48, 1), (64, 90, 2), (9, 30, 3)])
def test_combine_is_associative(M, N, seed):
    """The ranks fold the earlier shards' aggregates with a log-depth scan (pbwtgpu_shard_finish) and the windows of one
    GPU with a Hillis-Steele scan (windows.cu); both rely on the combine operator being associative and on the fold
    of all nodes being the state of the sequential sweep."""
    rng = np.random.default_rng(seed)
    hap = (rng.random((M, N)) < 0.25).astype(np.uint8)
    hap[: M // 3] = hap[0]
    a_start = rng.permutation(M).astype(np.int32)
    x = O.pack_panel(hap)
    cuts = [0, N // 4, N // 2, (3 * N) // 4, N]
    nodes = [aggregate_node(x[cuts[i]:cuts[i + 1]], M, cuts[i + 1] - cuts[i], cuts[i], a_start) for i in range(4)]
    A, B, C, D = nodes
    left = combine(combine(combine(A, B, M), C, M), D, M)
    tree = combine(combine(A, B, M), combine(C, D, M), M)
    right = combine(A, combine(B, combine(C, D, M), M), M)
    assert (left == tree).all() and (left == right).all()
    _, a_ref, d_ref = O.port().build(x, M, N, a_start, with_d=True)
    assert (split(left, M)[0] == a_ref).all()
    assert (split(left, M)[1][1:M] == d_ref[1:M]).all()
