#!/usr/bin/env python
"""bench.py -- headline benchmark of the PBWT build + maxWithin hot path (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one pass of the hot path over one synthetic panel: build (a[], d[], pack3 stream) +
matchMaximalWithin over M x N haplotype-sites.  metric = haplotype-sites/s.
  parity gate : before anything is timed, the GPU's build + maxWithin of the first sites of the workload must hash to
          the same values as the reference's (the cpu_baseline leg computes them); on a mismatch no value is printed.
  value : panel resident in HBM, CUDA events on the launching stream around each step.
  e2e   : the same step through the C-ABI call the host shim makes (pbwtgpu_build on host bit columns; outputs land in
          host buffers owned by the library), host<->device copies inside the timed region.
  configs : BASELINE.json configs[2..4] (100k x 200k build + pack3; 1M x 100k build with d; 1,000 queries vs a 100k
          panel), each with its own roofline fraction and CPU baseline; at N > 1 config 4 and the query sweep run
          site-sharded over the N GPUs.
  roofline / cpu_baseline: see DESIGN.md section "Measurement".
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "haplotype-sites/sec build+maxWithin"
UNIT = "haplotype-sites/s"
M_DEFAULT, N_DEFAULT = 10000, 100000  # BASELINE.json configs[1]
SEED, L_RECOMB, MU = 0x5EED0002, 2000, 2.0
BYTES_AD, BYTES_A = 16.125, 8.125  # algorithmic bytes per haplotype-site with / without d
L2_NOTE = ("GPU arm: 256 MiB device memset between timed steps, per-step CUDA events on the launching stream; "
           "CPU arm: one thread, nothing to flush")


def workload_name(M, N):
    return f"synthetic mosaic panel {M} haplotypes x {N} biallelic sites, build + maxWithin"


def base_config(M, n_total, world, nyz_all, nmatch_all):
    """The workload description, identical in the GPU arm and the reference arm."""
    return {"workload": workload_name(M, n_total), "haplotypes": M, "sites": n_total, "sites_per_gpu": n_total // world,
            "generator": f"seed {SEED:#x}, L={L_RECOMB}, mu={MU}", "l2": L2_NOTE,
            "matches_per_haplotype": nmatch_all / M, "pack3_bytes": nyz_all}


def host_info():
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return {"cpu_model": model, "nproc": os.cpu_count(), "nproc_allowed": len(os.sched_getaffinity(0))}


class PinnedToOneCore:
    """taskset for the CPU legs: the calling thread is pinned to one core (the reference is single-threaded)."""

    def __enter__(self):
        self.saved = os.sched_getaffinity(0)
        self.core = max(self.saved)
        os.sched_setaffinity(0, {self.core})
        return self

    def __exit__(self, *a):
        os.sched_setaffinity(0, self.saved)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], False
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def oracle():
    """(implementation, kind): the compiled reference when oracle/_ref exists, else the pinned port.  Only the
    cpu_baseline / parity-gate leg and --impl reference come here; nothing on the timed GPU path does."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O

    ref = O.reference()
    return (ref, "reference", O) if ref is not None else (O.port(), "port", O)


def time_reference(xbits, M, n_cpu, n_workload):
    """The reference's own single-threaded CPU path on the first n_cpu sites, pinned to one core; also returns the
    checksums of what it produced (oracle/oracle_api.h) for the parity gate."""
    imp, kind, _ = oracle()
    with PinnedToOneCore() as pin:
        tb, tm, nyz, nmatch, ck = imp.time_build_maxwithin_ck(xbits[:n_cpu], M, n_cpu)
    what = "all" if n_cpu == n_workload else "first"
    return {"value": M * n_cpu / (tb + tm), "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"{what} {n_cpu} of the workload's {n_workload} sites, all {M} haplotypes: build {tb:.2f}s + maxWithin {tm:.2f}s, "
                      f"{nmatch} matches (checksumming callback), single thread pinned to core {pin.core} (sched_setaffinity)",
            "host": host_info(), "seconds": tb + tm, "nyz": nyz, "nmatch": nmatch, "checksums": ck}


def gpu_checksums(O, out):
    return (O.checksum_join([O.checksum_parts(O.mix_bytes(out["yz"]))]) +
            O.checksum_join([O.checksum_parts(O.mix_records(out["matches"]))]))


def peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def reference_arm(args, M, N):
    """--impl reference: the unmodified reference (oracle/_ref) on this box's host cores, same workload and metric."""
    imp, kind, O = oracle()
    # whole workload per step unless that would take more than a few minutes in total
    n_cpu = args.cpu_sites or N
    x = O.generate_panel(M, n_cpu, SEED, L_RECOMB, MU)
    first = time_reference(x, M, n_cpu, N)
    reps = args.warmup + args.steps
    if not args.cpu_sites and first["seconds"] * reps > 300.0:  # keep the run within a few minutes: bounded sample
        n_cpu = max(2000, int(N * 300.0 / (first["seconds"] * reps)) // 1000 * 1000)
        first = time_reference(x, M, n_cpu, N)
    times, cb = [], first
    for i in range(reps):
        cb = time_reference(x, M, n_cpu, N)
        if i >= args.warmup:
            times.append(cb["seconds"])
    t = float(np.mean(times)) if times else cb["seconds"]
    val = M * n_cpu / t
    for k in ("seconds", "checksums"):
        cb.pop(k)
    nyz, nmatch = cb.pop("nyz"), cb.pop("nmatch")
    cb["value"] = val
    scale = N / n_cpu  # (only != 1 for a bounded sample)
    cfg = base_config(M, N, 1, int(nyz * scale), int(nmatch * scale))
    if n_cpu != N:
        cfg["step"] = f"bounded sample: first {n_cpu} sites (pack3_bytes / matches extrapolated)"
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                      "config": cfg, "cpu_baseline": cb,
                      "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ---------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[2..4] (single GPU unless stated)
# ---------------------------------------------------------------------------------------------------------------
def timed_ms(torch, stream, flush, fn, warmup, steps):
    """mean device time of fn() in ms (CUDA events on the launching stream, L2 flushed between runs)"""
    for _ in range(warmup):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush.zero_()
        a.record(stream)
        fn()
        b.record(stream)
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in ev]))


def roofline_of(algo_bytes, ms, peak):
    ach = algo_bytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak}


def run_config_build(P, torch, eng, stream, flush, name, M, N, seed, L, mu, flags, with_d, peak, n_cpu, cpu):
    """configs[2] / configs[3]: device-resident build of a wide panel; CPU baseline on the first n_cpu sites."""
    panel = eng.panel_synthetic(M, N, seed=seed, L=L, mu=mu)
    eng.synchronize()
    eng.last_timings()
    sizes = {}

    def step():
        r = eng.build_device(panel, flags)
        sizes["nyz"] = r.nyz
        r.free()

    ms = timed_ms(torch, stream, flush, step, 3, 2)
    kt = eng.last_timings()
    per_hs = BYTES_AD if with_d else BYTES_A
    out = {"workload": name, "haplotypes": M, "sites": N, "ms": ms, "value": M * N / (ms * 1e-3), "unit": UNIT,
           "pack3_bytes": sizes["nyz"], "algorithmic_bytes_per_hap_site": per_hs,
           "roofline": roofline_of(per_hs * M * N + sizes["nyz"], ms, peak),
           "kernel_ms_per_step": {k: v[0] / 5 for k, v in kt.items() if v[0] > 0}}
    if cpu:
        imp, kind, O = oracle()
        x = panel.to_host(0, n_cpu)
        sub = eng.panel_from_host(x, M)
        r = eng.build_device(sub, flags)
        got = r.fetch()
        r.free()
        sub.free()
        with PinnedToOneCore():
            tb, nyz, ck = imp.time_build(x, M, n_cpu, with_d)
        mine = O.checksum_join([O.checksum_parts(O.mix_bytes(got["yz"]))])
        out["cpu_baseline"] = {"value": M * n_cpu / tb, "unit": UNIT, "cores": 1, "kind": kind,
                               "sample": f"first {n_cpu} of {N} sites (extrapolated linearly), {'a+d' if with_d else 'a only'}: {tb:.2f}s"}
        out["parity"] = "ok" if (mine == ck and nyz == got["yz"].size) else "MISMATCH"
    panel.free()
    return out


def run_config_query(P, torch, eng, stream, flush, M, Q, N, peak, n_cpu, cpu):
    """configs[4]: 1,000 queries (the last Q haplotypes of an (M+Q)-panel) against the other M, matchSequencesSweep."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fullsize_checks import split_panel

    both = eng.panel_synthetic(M + Q, N, seed=0x5EED0005, L=1000, mu=8.0)
    x = both.to_host()
    both.free()
    px, qx = split_panel(x, M, Q)
    del x
    panel, queries = eng.panel_from_host(px, M), eng.panel_from_host(qx, Q)
    rp = eng.build_device(panel, P.WANT_YZ); pyz = rp.fetch()["yz"]; rp.free()
    rq = eng.build_device(queries, P.WANT_YZ); qyz = rq.fetch()["yz"]; rq.free()
    panel.free()
    queries.free()
    eng.synchronize()
    eng.last_timings()
    info = {}

    def step():
        rec, nom = eng.match_sweep(pyz, M, N, qyz, Q)
        info["nrec"], info["nnom"] = len(rec), len(nom)

    step()
    eng.last_timings()
    t0 = time.time()
    step()
    wall = time.time() - t0
    kt = eng.last_timings()
    dev_ms = sum(v[0] for v in kt.values())
    out = {"workload": f"matchSequencesSweep: {Q} queries vs {M} haplotypes x {N} sites (N assumed: BASELINE gives none)",
           "haplotypes": M, "queries": Q, "sites": N, "ms": dev_ms, "value": M * N / (dev_ms * 1e-3), "unit": "panel " + UNIT,
           "wall_ms_host_streams_in_records_out": wall * 1e3, "records": info["nrec"], "no_match_events": info["nnom"],
           "algorithmic_bytes_per_hap_site": BYTES_AD,
           "roofline": roofline_of(BYTES_AD * M * N + 16.0 * info["nrec"], dev_ms, peak),
           "kernel_ms_per_step": {k: v[0] for k, v in kt.items() if v[0] > 0}}
    if cpu:
        imp, kind, O = oracle()
        pp, qq = eng.panel_from_host(px[:n_cpu], M), eng.panel_from_host(qx[:n_cpu], Q)
        r = eng.build_device(pp, P.WANT_YZ); pyz5 = r.fetch()["yz"]; r.free()
        r = eng.build_device(qq, P.WANT_YZ); qyz5 = r.fetch()["yz"]; r.free()
        pp.free()
        qq.free()
        rec, nom = eng.match_sweep(pyz5, M, n_cpu, qyz5, Q)
        with PinnedToOneCore():
            t0 = time.time()
            want, wnom = imp.match_sweep(pyz5, M, n_cpu, qyz5, Q)
            tc = (time.time() - t0) / 2  # (the wrapper runs it twice: count, then fill)
        out["cpu_baseline"] = {"value": M * n_cpu / tc, "unit": "panel " + UNIT, "cores": 1, "kind": kind,
                               "sample": f"first {n_cpu} of {N} sites (extrapolated linearly): matchSequencesSweep {tc:.2f}s, {len(want)} records"}
        out["parity"] = "ok" if (np.array_equal(rec, want) and np.array_equal(nom, wnom)) else "MISMATCH"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--haps", type=int, default=M_DEFAULT)
    ap.add_argument("--sites", type=int, default=N_DEFAULT)
    ap.add_argument("--cpu-sites", type=int, default=0, help="sites timed on the CPU baseline (0 = automatic)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU legs (and with them the parity gate)")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE configs 3-5 (the headline line only)")
    ap.add_argument("--no-cli", action="store_true", help="skip the command-line wall times")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    M, N = args.haps, args.sites
    warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, M, N)
        return

    import torch
    import torch.distributed as dist

    import pbwt_b200 as P

    numa_node = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    eng = P.Engine(local_rank)
    # an explicit (non-default) stream: the engine's kernels, the CUDA events and the NCCL all-gather are all ordered on it
    # (the legacy default stream has handle 0, which pbwtgpu_set_stream reads as "use the context's own stream")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    eng.profile(True)
    flags = P.WANT_YZ | P.WANT_MAXWITHIN
    peak, peak_src = peak_hbm()

    # ---- parity gate (rank 0): GPU vs the reference on the first sites of the workload, before any timing ----
    cpu = None
    gate = {"checked": False}
    if rank == 0 and not args.no_cpu_baseline:
        n_cpu = args.cpu_sites or min(N, 20000)
        head = eng.panel_synthetic(M, n_cpu, seed=SEED, L=L_RECOMB, mu=MU)
        r = eng.build_device(head, flags)
        got = r.fetch()
        grouped = r.fetch_groups()  # the form the e2e path brings to the host: must expand to the very same records
        r.free()
        x_head = head.to_host()
        head.free()
        cpu = time_reference(x_head, M, n_cpu, N * world)
        _, _, O = oracle()
        mine = gpu_checksums(O, got)
        gate = {"checked": True, "sites": n_cpu, "pack3_bytes": int(got["yz"].size), "records": int(len(got["matches"])),
                "checksums_gpu": mine, "checksums_cpu": cpu["checksums"],
                "grouped_words": int(grouped.size),
                "ok": (mine == cpu["checksums"] and got["yz"].size == cpu["nyz"] and len(got["matches"]) == cpu["nmatch"]
                       and np.array_equal(P.expand_groups(grouped), got["matches"]))}
        if not gate["ok"]:
            print(json.dumps({"error": "parity gate failed: the GPU's build + maxWithin of the first sites differs from the "
                                       "reference's; no value is reported", "parity_gate": gate}))
            sys.exit(3)
        for k in ("seconds", "checksums", "nyz", "nmatch"):
            cpu.pop(k)

    # N GPUs: the SITES of one panel of world*N sites are partitioned across the ranks (weak scaling: N sites per
    # GPU); the only exchange is one NCCL all-gather of per-shard aggregates (include/pbwtgpu.h, multi-GPU)
    n_total = N * world
    panel = eng.panel_synthetic(M, N, seed=SEED, L=L_RECOMB, mu=MU, first_site=rank * N)
    eng.synchronize()
    eng.last_timings()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    agg_ints = eng.shard_aggregate_ints(M)
    agg = torch.zeros(agg_ints, dtype=torch.int32, device="cuda")
    aggs = torch.zeros((world, agg_ints), dtype=torch.int32, device="cuda")

    def device_step(pnl):
        """One build + maxWithin step over this rank's sites; returns the Result (device resident)."""
        if world == 1:
            return eng.build_device(pnl, flags)
        sh = eng.shard_prepare(pnl, rank * N, n_total, flags)
        sh.export(agg.data_ptr())
        dist.all_gather_into_tensor(aggs, agg)
        res = sh.finish(aggs.data_ptr(), rank, flags)
        sh.free()
        return res

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- multi-GPU parity (N > 1): the shards' outputs, concatenated in site order, against ONE GPU building the
    # whole panel (rank 0), by checksums of pack3 bytes and records, plus aFend / d of the last shard ----
    mg_parity = None
    if world > 1:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O

        r = device_step(panel)
        out = r.fetch()
        r.free()
        # (uint64 sums travel as int64 bit patterns)
        parts = torch.from_numpy(np.array([[out["yz"].size, *O.checksum_parts(O.mix_bytes(out["yz"]))[1:]],
                                           [len(out["matches"]), *O.checksum_parts(O.mix_records(out["matches"]))[1:]]],
                                          dtype=np.uint64).view(np.int64)).cuda()
        allparts = torch.zeros((world, 2, 3), dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(allparts, parts)
        ends = torch.from_numpy(np.concatenate([out["a_end"], out["d_end"]])).cuda()
        dist.broadcast(ends, src=world - 1)
        if rank == 0:
            ap_ = allparts.cpu().numpy().view(np.uint64)
            joined = (O.checksum_join([tuple(int(v) for v in ap_[w, 0]) for w in range(world)]) +
                      O.checksum_join([tuple(int(v) for v in ap_[w, 1]) for w in range(world)]))
            whole = eng.panel_synthetic(M, n_total, seed=SEED, L=L_RECOMB, mu=MU)
            r1 = eng.build_device(whole, flags)
            o1 = r1.fetch()
            r1.free()
            whole.free()
            single = gpu_checksums(O, o1)
            e = ends.cpu().numpy()
            ok = (joined == single and int(ap_[:, 0, 0].sum()) == o1["yz"].size and int(ap_[:, 1, 0].sum()) == len(o1["matches"])
                  and np.array_equal(e[:M], o1["a_end"]) and np.array_equal(e[M:], o1["d_end"]))
            mg_parity = {"ok": bool(ok), "what": f"{world} shards concatenated vs one GPU building all {n_total} sites: checksums of "
                                                 f"{o1['yz'].size} pack3 bytes and {len(o1['matches'])} records, aFend, d"}
            if not ok:
                print(json.dumps({"error": "multi-GPU parity failed: the shards' outputs differ from the single-GPU build",
                                  "parity": mg_parity}))
        flag = torch.tensor([1 if (rank != 0 or mg_parity["ok"]) else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag) == 0:
            dist.destroy_process_group()
            sys.exit(3)

    # ---- device-resident steps -------------------------------------------------------------------
    nyz = nmatch = 0
    for _ in range(warmup):
        r = device_step(panel)
        nyz, nmatch = r.nyz, r.nmatch
        r.free()
    eng.synchronize()
    eng.last_timings()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    with ClockSampler(local_rank) as clocks:
        t_wall0 = time.time()
        for s in range(args.steps):
            flush.zero_()  # evict L2 between timed steps (not inside the step's events)
            ev[s][0].record(stream)
            r = device_step(panel)
            ev[s][1].record(stream)
            r.free()
        barrier()
        t_wall = time.time() - t_wall0
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    ms_per_step = float(np.mean(ms_steps))
    kt = eng.last_timings()
    launches = sum(v[1] for v in kt.values())
    # one step's sweep work: pass 1 (window nodes: transpose + window sort), the combine rounds and pass 2 (both the
    # sweep kernel); the roofline is quoted on their summed duration
    sweep_ms = kt["sweep"][0] + kt["prepass"][0] + kt["combine"][0]
    sweep_launches = kt["sweep"][1] + kt["prepass"][1] + kt["combine"][1]

    # ---- end to end through the C ABI as the host shim calls it ------------------------------------
    # bit columns in a page-locked buffer of the library's pool (the shim assembles its reader batches in one:
    # pbwtgpu_host_alloc), ONE pbwtgpu_build call per step (H2D, kernels, D2H inside), outputs in library-owned host
    # buffers that are read in place and given back with pbwtgpu_free -- exactly the calls of pbwt_gpu_shim.c
    wpc = P.words_per_col(M)
    xbuf = eng.host_alloc(N * wpc * 4)
    x_np = xbuf.view(np.uint32).reshape(N, wpc)
    x_np[:] = panel.to_host()
    a_end, d_end = np.zeros(M, np.int32), np.zeros(M + 1, np.int32)
    h2d = x_np.nbytes
    d2h = nyz + 4 * M + 4 * (M + 1) + 16 * nmatch  # (replaced below by what the grouped stream really is)
    pin = None
    if world > 1:  # per-rank pinned output buffers for the device-resident calls of the sharded protocol
        # (the matches come down in the grouped form, as at N = 1: {ai, start, end, n} + n x bi, at most 5 words per record)
        pin = {"yz": torch.empty(max(nyz, 1), dtype=torch.uint8).pin_memory().numpy(),
               "a_end": torch.empty(M, dtype=torch.int32).pin_memory().numpy(),
               "d_end": torch.empty(M + 1, dtype=torch.int32).pin_memory().numpy(),
               "groups": None}

    def group_words(e, res):  # the grouped stream is built on the device by the context that owns the result
        nw, ng = ctypes.c_int64(0), ctypes.c_int64(0)
        e._check(e.lib.pbwtgpu_result_group(e.ctx, res.handle, ctypes.byref(nw), ctypes.byref(ng)))
        return nw.value

    def fetch_with(e, res, out):  # pack3 stream, aFend, d and the grouped matches, through context e
        e._check(e.lib.pbwtgpu_result_fetch(e.ctx, res.handle, out["yz"].ctypes.data, out["a_end"].ctypes.data, out["d_end"].ctypes.data, None))
        e._check(e.lib.pbwtgpu_result_fetch_groups(e.ctx, res.handle, out["groups"].ctypes.data))

    def abi_step(e, a_out, d_out):
        # the matches come as the grouped stream ({ai, start, end, n} + n x bi per run of callbacks that share ai, start,
        # end: include/pbwtgpu.h), which the shim expands while it calls report(); the host reads it where the library left it
        yzp, ny, gp, nw = e.build_grouped_raw(x_np, M, N, flags, a_out, d_out)
        last = np.ctypeslib.as_array(ctypes.cast(gp, ctypes.POINTER(ctypes.c_int32)), shape=(nw,))[-1] if nw else 0
        e.free_raw(yzp, gp)
        return ny, nw, int(last)

    def abi_step_full_records(e, a_out, d_out):  # the same call with 16-byte records (pbwtgpu_build), for comparison
        yzp, ny, mtp, nm = e.build_raw(x_np, M, N, flags, a_out, d_out)
        e.free_raw(yzp, mtp)

    def e2e_step():
        if world == 1:
            return abi_step(eng, a_end, d_end)
        pnl = eng.panel_from_host(x_np, M)  # N > 1: the sharded protocol needs the device-resident calls
        res = device_step(pnl)
        nw = group_words(eng, res)
        if pin["groups"] is None:
            pin["groups"] = torch.empty(max(nw, 1), dtype=torch.int32).pin_memory().numpy()
        fetch_with(eng, res, pin)
        res.free()
        pnl.free()
        return 0, nw, 0

    first = e2e_step()
    e2e_step()
    d2h = nyz + 4 * M + 4 * (M + 1) + 4 * first[1]
    barrier()
    t0 = time.time()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    t_e2e_serial = (time.time() - t0) / args.steps
    t_e2e, in_flight = t_e2e_serial, 1
    shim_seq_ms = full_records_ms = None

    pipelined_ms = None
    if world > 1:
        # N > 1: a three-stage pipeline per rank -- a host thread with its own context uploads the next step's columns, the
        # main thread runs the sharded build (the only one that touches NCCL), a third thread with its own context downloads
        # the previous step's results -- so that each step's PCIe copies overlap its neighbours' kernels.  Every step still
        # does its own H2D, kernels + all-gather and D2H inside the timed region; device blocks go back to the context
        # that allocated them.
        import queue

        eng_up, eng_down = P.Engine(local_rank), P.Engine(local_rank)
        # more waiting host threads than cores (3 per rank): the library's waits sleep on an event instead of spinning
        blocking = 3 * world > (os.cpu_count() or 1)
        for e in (eng, eng_up, eng_down):
            e.set_option("blocking_sync", 1 if blocking else 0)
        pins = [pin, {k: torch.from_numpy(v.view(np.uint8)).clone().pin_memory().numpy().view(v.dtype).reshape(v.shape) for k, v in pin.items()}]

        def pipeline(steps):
            q_up, q_down, q_free_panels, q_free_results = queue.Queue(2), queue.Queue(2), queue.Queue(), queue.Queue()

            def uploader():
                for _ in range(steps):
                    q_up.put(eng_up.panel_from_host(x_np, M))
                    while not q_free_panels.empty():
                        q_free_panels.get().free()
                q_up.put(None)

            def downloader():
                i = 0
                while True:
                    item = q_down.get()
                    if item is None:
                        break
                    res, pnl = item
                    fetch_with(eng_down, res, pins[i & 1])
                    i += 1
                    q_free_results.put(res)
                    q_free_panels.put(pnl)

            tu, td = threading.Thread(target=uploader), threading.Thread(target=downloader)
            tu.start()
            td.start()
            while True:
                pnl = q_up.get()
                if pnl is None:
                    break
                res = device_step(pnl)
                group_words(eng, res)
                q_down.put((res, pnl))
                while not q_free_results.empty():
                    q_free_results.get().free()
            q_down.put(None)
            tu.join()
            td.join()
            while not q_free_results.empty():
                q_free_results.get().free()
            while not q_free_panels.empty():
                q_free_panels.get().free()

        pipeline(3)
        barrier()
        t0 = time.time()
        pipeline(args.steps)
        barrier()
        pipelined_ms = (time.time() - t0) / args.steps * 1e3
        eng_up.close()
        eng_down.close()
        eng.set_option("blocking_sync", 0)
        if pipelined_ms * 1e-3 < t_e2e:
            t_e2e, in_flight = pipelined_ms * 1e-3, 3

    if world == 1:
        # Two builds in flight: two contexts, each with its own stream and host thread, take the steps alternately, so
        # that one step's PCIe copies overlap the other's kernels.  Every step is still one complete pbwtgpu_build call.
        engs = [P.Engine(local_rank), P.Engine(local_rank)]
        outs = [(np.zeros(M, np.int32), np.zeros(M + 1, np.int32)) for _ in engs]

        def worker(w, steps):
            for _ in range(steps):
                abi_step(engs[w], *outs[w])

        for w in (0, 1):  # warm both contexts (their device-memory caches, kernel attributes)
            worker(w, 2)
        barrier()
        counts = [(args.steps + 1) // 2, args.steps // 2]
        threads = [threading.Thread(target=worker, args=(w, counts[w])) for w in (0, 1) if counts[w]]
        t0 = time.time()
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        barrier()
        t_pipe = (time.time() - t0) / args.steps
        for e in engs:
            e.close()
        if t_pipe < t_e2e:
            t_e2e, in_flight = t_pipe, 2
        # the two-call sequence of `pbwt_gpu -readMacs ... -maxWithin`: the readers' batch goes through pbwtgpu_build
        # (pack3 stream + aFend back), then matchMaximalWithin sends the stored stream to pbwtgpu_max_within
        def shim_sequence():
            yzp, ny, _, _ = eng.build_raw(x_np, M, N, P.WANT_YZ, a_end, None)
            gp, nw = eng.max_within_grouped_raw(yzp, ny, M, N)
            eng.free_raw(yzp, gp)

        shim_sequence()
        barrier()
        t0 = time.time()
        for _ in range(args.steps):
            shim_sequence()
        barrier()
        shim_seq_ms = (time.time() - t0) / args.steps * 1e3
        abi_step_full_records(eng, a_end, d_end)
        barrier()
        t0 = time.time()
        for _ in range(args.steps):
            abi_step_full_records(eng, a_end, d_end)
        barrier()
        full_records_ms = (time.time() - t0) / args.steps * 1e3
    eng.host_free(xbuf)

    # ---- command-line wall times (reported separately; text output dominates) ------------
    cli = None
    if rank == 0 and world == 1 and not args.no_cli and (M, N) == (M_DEFAULT, N_DEFAULT):
        cli = cli_wall_times(eng, P, panel, M, N)

    # ---- BASELINE configs 3-5 -----------------------------------------------------------------------
    configs = None
    if not args.no_configs and (M, N) == (M_DEFAULT, N_DEFAULT):
        panel.free()
        panel = None
        if world == 1:
            want_cpu = not args.no_cpu_baseline
            configs = {
                "c3_build_pack3_100k_x_200k": run_config_build(
                    P, torch, eng, stream, flush, "synthetic mosaic panel 100000 x 200000, build + pack3 (a only)", 100000, 200000,
                    0x5EED0003, 1000, 8.0, P.WANT_YZ, False, peak, 5000, want_cpu),
                "c4_build_ad_1m_x_100k_one_gpu": run_config_build(
                    P, torch, eng, stream, flush, "synthetic mosaic panel 1000000 x 100000, build with d + pack3, ONE GPU", 1000000, 100000,
                    0x5EED0004, 1000, 8.0, P.WANT_YZ | P.WANT_D, True, peak, 5000, want_cpu),
                "c5_match_sweep_1000_vs_100k_x_100k": run_config_query(P, torch, eng, stream, flush, 100000, 1000, 100000, peak, 5000, want_cpu),
            }
        else:
            configs = {"c4_build_ad_1m_x_100k_site_sharded": run_config4_sharded(P, torch, dist, eng, stream, flush, rank, world, peak),
                       "c5_match_sweep_1000_vs_100k_x_100k_site_sharded": run_config5_sharded(P, torch, dist, eng, stream, rank, world, peak)}
    if panel is not None:
        panel.free()

    # ---- reduce over ranks (max time) --------------------------------------------------------------
    t_step = torch.tensor([ms_per_step, t_e2e * 1e3], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_step, op=dist.ReduceOp.MAX)
    ms_per_step, ms_e2e = float(t_step[0]), float(t_step[1])
    total_hs = float(M) * N * world
    stats = torch.tensor([float(nyz), float(nmatch), float(h2d), float(d2h), float(launches)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats)
    nyz_all, nmatch_all, h2d_all, d2h_all, launches_all = [int(v) for v in stats.tolist()]

    if rank == 0:
        algo_bytes = BYTES_AD * M * N + nyz + 16.0 * nmatch  # rank 0's shard; the roofline is per GPU
        # one unit of algorithmic work (a step) needs ALL sweep-kernel launches of that step (pass 1 + combine
        # rounds + pass 2), so the denominator is their summed duration per step, not one launch's
        sweep_ms_per_step = sweep_ms / args.steps
        achieved = algo_bytes / (sweep_ms_per_step * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    # dram__bytes_read+write per step from the ncu --set full captures under profiles/ (pass 2, the dominant
                    # launch, plus window sort and transpose); only valid for this config
                    "traffic": TRAFFIC_C2 if (M, N) == (M_DEFAULT, N_DEFAULT) else None, "kernel": "sweep_smem_kernel",
                    "frac_of_whole_step": algo_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                    "kernel_ms_per_step": sweep_ms_per_step, "kernel_launches_per_step": sweep_launches / args.steps,
                    "kernel_ms_per_step_by_role": {"pass1": kt["prepass"][0] / args.steps, "combine": kt["combine"][0] / args.steps,
                                                   "pass2": kt["sweep"][0] / args.steps},
                    "algorithmic_bytes_per_step": algo_bytes, "peak_source": peak_src,
                    "note": "algorithmic bytes (16.125 B per haplotype-site + emitted bytes) over the summed duration of the "
                            "launches that do a step's sweep work: pass 1 (window nodes by transpose + radix sort of window "
                            "keys), combine rounds and pass 2 (sweep_smem_kernel, the dominant launch); a[]/d[] stay resident "
                            "in shared memory, so DRAM traffic is far below the algorithmic bytes"}
        cfg = base_config(M, n_total, world, nyz_all, nmatch_all)
        cfg["parallelism"] = (f"sites partitioned across {world} GPUs, one NCCL all-gather of per-shard aggregates "
                              f"({agg_ints * 4} B per rank) per step") if world > 1 else "single GPU"
        out = {
            "metric": METRIC, "value": total_hs / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic", "config": cfg,
            "e2e": {"value": total_hs / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h_all,
                    "ms_per_step": ms_e2e, "builds_in_flight": in_flight, "ms_per_step_one_in_flight": t_e2e_serial * 1e3,
                    "ms_per_step_pipelined_rank0": pipelined_ms, "host_cores": os.cpu_count(), "numa_node_rank0": numa_node,
                    "ms_per_step_shim_two_call_sequence": shim_seq_ms,
                    "ms_per_step_one_in_flight_16_byte_records": full_records_ms,
                    "note": ("one pbwtgpu_build_grouped(host bit columns in a pbwtgpu_host_alloc buffer, WANT_YZ|WANT_MAXWITHIN) per step, "
                             "the entry point the C host shim's match path uses: H2D, kernels and D2H of pack3 stream, aFend, d and ALL "
                             "match records -- as the grouped stream ({ai, start, end, n} + n x bi per run of callbacks that share "
                             "ai, start, end; the shim expands it while it calls report(); the parity gate checks that it expands "
                             "to exactly the records) -- into library-owned host buffers, freed with pbwtgpu_free; with 2 builds in "
                             "flight two contexts (streams, host threads) alternate steps; the shim's two-call sequence "
                             "(pbwtgpu_build, then pbwtgpu_max_within_grouped on the stored stream) and the same call with 16-byte "
                             "records are timed beside it") if world == 1 else
                            "per rank: pinned host columns -> pbwtgpu_panel_from_host -> shard protocol (one NCCL all-gather) -> "
                            "pbwtgpu_result_group + pbwtgpu_result_fetch / _fetch_groups of pack3 stream, aFend, d and all match "
                            "records in the grouped form, every step; builds_in_flight 3 = "
                            "a three-stage pipeline per rank (upload thread, build thread, download thread, each with its own "
                            "context and stream), 1 = strictly one step after the other"},
            "gpu_launches": launches_all,
            "kernel_ms": {k: v[0] for k, v in kt.items() if v[0] > 0},
            "clocks": clocks.summary(),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity_gate": gate,
            "wall_s_timed_region": t_wall,
        }
        if mg_parity:
            out["parity"] = mg_parity
        if cli:
            out["cli"] = cli
        if configs:
            out["configs"] = configs
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum of the dominant launch of a config-2 step, pass 2 of sweep_smem_kernel
# (ncu --set full, profiles/r02f_c2_pass2_ncu_full_summary.txt: 139.3 MB read -- the panel columns once -- and 255.2 MB
# written -- records and run lists).  The step's other sweep launches: pass 1's window sort 182 MB (r01c_*), the eight
# combine launches 26 MB each (r02f_c2_combine_*).
TRAFFIC_C2 = 394.5e6


def bind_to_gpu_numa_node(torch, local_rank):
    """N > 1: run this rank on the cores of its GPU's NUMA node, so that the page-locked host buffers it allocates
    (first touch) and the threads that fill them sit next to the GPU's PCIe root -- eight ranks moving 1.7 GB per step
    through one socket's memory and the inter-socket link is what the end-to-end number at N = 8 waits for.  Returns the
    node, or None when the topology cannot be read (nothing is changed then)."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if len(cpus) < 2:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def cli_wall_times(eng, P, panel, M, N):
    """`pbwt_gpu -read X.pbwt -maxWithin > /dev/null` next to the reference's CLI on the same file (one run each)."""
    gpu_cli = os.path.join(ROOT, "pbwt_b200", "host", "_build", "pbwt_gpu")
    ref_cli = os.path.join(ROOT, "oracle", "_ref", "pbwt_ref")
    if not (os.path.exists(gpu_cli) and os.path.exists(ref_cli)):
        return {"unavailable": "pbwt_gpu / pbwt_ref were not prebuilt (they need the reference's sources)"}
    r = eng.build_device(panel, P.WANT_YZ)
    o = r.fetch()
    r.free()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "c2.pbwt")
        with open(path, "wb") as fh:  # .pbwt v3, pbwtIO.c:33-57
            fh.write(b"PBW3" + np.array([M, N], "<i4").tobytes() + np.arange(M, dtype="<i4").tobytes() + o["a_end"].astype("<i4").tobytes() +
                     np.array([o["yz"].size], "<i8").tobytes() + b"    " + o["yz"].tobytes())
        env = dict(os.environ, PBWTGPU_SHIM_TIMING="1")  # the shim says on stderr where its time went
        for name, exe in (("pbwt_gpu", gpu_cli), ("pbwt_gpu_second_run", gpu_cli), ("pbwt_ref", ref_cli)):
            t0 = time.time()
            with open(os.devnull, "wb") as null:
                pr = subprocess.run([exe, "-read", path, "-maxWithin"], stdout=null, stderr=subprocess.PIPE, env=env)
            out[name + "_s"] = time.time() - t0
            out[name + "_rc"] = pr.returncode
            phases = {}
            for ln in pr.stderr.decode(errors="replace").splitlines():
                if ln.startswith("pbwtgpu shim timing: "):
                    what, sec = ln[len("pbwtgpu shim timing: "):].rsplit(" ", 2)[0:2]
                    phases[what] = float(sec)
            if phases:
                out[name + "_phases_s"] = phases
    out["command"] = ("<cli> -read c2.pbwt -maxWithin > /dev/null (process start, file read, 18.5 M MATCH lines formatted); pbwt_gpu runs while "
                      "this process keeps its own device context and buffers")
    return out


def run_config4_sharded(P, torch, dist, eng, stream, flush, rank, world, peak):
    """configs[3] at N > 1: the 1,000,000 x 100,000 panel with its SITES partitioned across the ranks (strong scaling),
    a + d + pack3; parity of the concatenated pack3 stream and of aFend / d against checksums is by rank 0."""
    M, N = 1000000, 100000
    n_loc = N // world
    first = rank * n_loc
    if rank == world - 1:
        n_loc = N - first
    flags = P.WANT_YZ | P.WANT_D
    panel = eng.panel_synthetic(M, n_loc, seed=0x5EED0004, L=1000, mu=8.0, first_site=first)
    agg_ints = eng.shard_aggregate_ints(M)
    agg = torch.zeros(agg_ints, dtype=torch.int32, device="cuda")
    aggs = torch.zeros((world, agg_ints), dtype=torch.int32, device="cuda")
    sizes = {}

    def step():
        sh = eng.shard_prepare(panel, first, N, flags)
        sh.export(agg.data_ptr())
        dist.all_gather_into_tensor(aggs, agg)
        res = sh.finish(aggs.data_ptr(), rank, flags)
        sh.free()
        sizes["nyz"] = res.nyz
        return res

    for _ in range(2):
        step().free()
    eng.synchronize()
    eng.last_timings()
    ms = []
    for _ in range(2):
        flush.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        res = step()
        b.record(stream)
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
        res.free()
    kt = eng.last_timings()
    t = torch.tensor([float(np.mean(ms))], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot = torch.tensor([float(sizes["nyz"])], device="cuda", dtype=torch.float64)
    dist.all_reduce(tot)
    # parity: pack3 bytes of all shards in site order, aFend and d of the last one, against ONE GPU building the panel
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O

    res = step()
    out = res.fetch()
    res.free()
    panel.free()
    parts = torch.from_numpy(np.array(O.checksum_parts(O.mix_bytes(out["yz"])), dtype=np.uint64).view(np.int64)).cuda()
    allparts = torch.zeros((world, 3), dtype=torch.int64, device="cuda")
    dist.all_gather_into_tensor(allparts, parts)
    ends = torch.from_numpy(np.concatenate([out["a_end"], out["d_end"]])).cuda()
    dist.broadcast(ends, src=world - 1)
    parity, one_gpu_ms = None, None
    if rank == 0:
        ap_ = allparts.cpu().numpy().view(np.uint64)
        joined = O.checksum_join([tuple(int(v) for v in ap_[w]) for w in range(world)])
        whole = eng.panel_synthetic(M, N, seed=0x5EED0004, L=1000, mu=8.0)
        eng.build_device(whole, flags).free()
        eng.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        r1 = eng.build_device(whole, flags)
        b.record(stream)
        torch.cuda.synchronize()
        one_gpu_ms = a.elapsed_time(b)
        o1 = r1.fetch()
        r1.free()
        whole.free()
        e = ends.cpu().numpy()
        ok = (joined == O.checksum_join([O.checksum_parts(O.mix_bytes(o1["yz"]))]) and int(ap_[:, 0].sum()) == o1["yz"].size
              and np.array_equal(e[:M], o1["a_end"]) and np.array_equal(e[M:], o1["d_end"]))
        parity = "ok" if ok else "MISMATCH"
    dist.barrier()
    ms_max = float(t[0])
    return {"parity_vs_one_gpu": parity, "one_gpu_ms_same_box": one_gpu_ms,
            "strong_scaling_efficiency": (one_gpu_ms / ms_max / world) if one_gpu_ms else None,"workload": f"synthetic mosaic panel {M} x {N}, build with d + pack3, sites partitioned across {world} GPUs (strong scaling)",
            "haplotypes": M, "sites": N, "n_gpus": world, "ms": ms_max, "value": M * N / (ms_max * 1e-3), "unit": UNIT,
            "pack3_bytes": int(tot[0]), "all_gather_bytes_per_rank": agg_ints * 4, "all_gather_bytes_total": agg_ints * 4 * world,
            "algorithmic_bytes_per_hap_site": BYTES_AD,
            "roofline_per_gpu": roofline_of((BYTES_AD * M * N + float(tot[0])) / world, ms_max, peak),
            "kernel_ms_per_step_rank0": {k: v[0] / 2 for k, v in kt.items() if v[0] > 0}}


def run_config5_sharded(P, torch, dist, eng, stream, rank, world, peak):
    """configs[4] at N > 1: matchSequencesSweep of 1,000 queries against 100,000 x 100,000 with the SITES partitioned
    across the ranks (include/pbwtgpu.h, pbwtgpu_qshard_*): two all-gathers of aggregates and the queries' state handed
    from rank to rank.  Every query is swept on every GPU's site range -- sharding the queries themselves would repeat
    the whole panel sweep on every GPU.  Parity: rank 0 runs the whole sweep on one GPU and compares the
    concatenated records and events by checksum."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from fullsize_checks import split_panel

    M, Q, N = 100000, 1000, 100000
    # the stored streams: built once on rank 0's GPU, broadcast as bytes (every rank reads the same two files in practice)
    if rank == 0:
        both = eng.panel_synthetic(M + Q, N, seed=0x5EED0005, L=1000, mu=8.0)
        x = both.to_host()
        both.free()
        px, qx = split_panel(x, M, Q)
        del x
        pp, qq = eng.panel_from_host(px, M), eng.panel_from_host(qx, Q)
        r = eng.build_device(pp, P.WANT_YZ); pyz = r.fetch()["yz"]; r.free()
        r = eng.build_device(qq, P.WANT_YZ); qyz = r.fetch()["yz"]; r.free()
        pp.free()
        qq.free()
        sizes = torch.tensor([pyz.size, qyz.size], dtype=torch.int64, device="cuda")
    else:
        sizes = torch.zeros(2, dtype=torch.int64, device="cuda")
    dist.broadcast(sizes, src=0)
    bufs = []
    for i, nbytes in enumerate(int(v) for v in sizes.tolist()):
        t = torch.from_numpy((pyz, qyz)[i].copy()).cuda() if rank == 0 else torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        dist.broadcast(t, src=0)
        bufs.append(t.cpu().numpy())
    pyz, qyz = bufs
    bounds = [N * r // world for r in range(world + 1)]
    qaggs = torch.zeros((world, eng.shard_aggregate_ints(Q)), dtype=torch.int32, device="cuda")
    paggs = torch.zeros((world, eng.shard_aggregate_ints(M)), dtype=torch.int32, device="cuda")
    qagg, pagg = torch.zeros_like(qaggs[0]), torch.zeros_like(paggs[0])
    s_in = torch.zeros((2, Q), dtype=torch.int32, device="cuda")
    s_out = torch.zeros((2, Q), dtype=torch.int32, device="cuda")

    def step():
        sh = eng.query_shard(pyz, M, N, qyz, Q, bounds[rank], bounds[rank + 1] - bounds[rank])
        sh.export_queries(qagg.data_ptr())
        dist.all_gather_into_tensor(qaggs, qagg)
        sh.prepare(qaggs.data_ptr(), rank, pagg.data_ptr())
        dist.all_gather_into_tensor(paggs, pagg)
        if rank > 0:
            dist.recv(s_in, src=rank - 1)
        sh.chain(paggs.data_ptr(), rank, s_in.data_ptr() if rank else None, s_out.data_ptr())
        if rank + 1 < world:
            dist.send(s_out, dst=rank + 1)
        out = sh.finish()
        sh.free()
        return out

    step()
    eng.synchronize()
    eng.last_timings()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    rec, nom = step()
    torch.cuda.synchronize()
    wall = time.time() - t0
    kt = eng.last_timings()
    t = torch.tensor([wall, sum(v[0] for v in kt.values()) * 1e-3], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    parts = torch.from_numpy(np.array([O.checksum_parts(O.mix_records(rec)), (len(nom), int(nom["site"].sum()) if len(nom) else 0, 0)],
                                      dtype=np.uint64).view(np.int64)).cuda()
    allparts = torch.zeros((world, 2, 3), dtype=torch.int64, device="cuda")
    dist.all_gather_into_tensor(allparts, parts)
    parity, one_gpu_ms = None, None
    if rank == 0:
        ap_ = allparts.cpu().numpy().view(np.uint64)
        joined = O.checksum_join([tuple(int(v) for v in ap_[w, 0]) for w in range(world)])
        eng.match_sweep(pyz, M, N, qyz, Q)
        eng.last_timings()
        rec1, nom1 = eng.match_sweep(pyz, M, N, qyz, Q)
        one_gpu_ms = sum(v[0] for v in eng.last_timings().values())
        ok = (joined == O.checksum_join([O.checksum_parts(O.mix_records(rec1))]) and int(ap_[:, 0, 0].sum()) == len(rec1)
              and int(ap_[:, 1, 0].sum()) == len(nom1) and int(ap_[:, 1, 1].sum()) == (int(nom1["site"].sum()) if len(nom1) else 0))
        parity = "ok" if ok else "MISMATCH"
    dist.barrier()
    wall_max, dev_max = float(t[0]), float(t[1])
    nrec = int(allparts[:, 0, 0].sum())
    return {"workload": f"matchSequencesSweep: {Q} queries vs {M} haplotypes x {N} sites, sites partitioned across {world} GPUs, every "
                        f"query on every GPU (strong scaling; N assumed: BASELINE gives none)",
            "haplotypes": M, "queries": Q, "sites": N, "n_gpus": world, "ms_wall_host_streams_in_records_out": wall_max * 1e3,
            "ms_kernels_max_over_ranks": dev_max * 1e3, "ms": dev_max * 1e3, "value": M * N / dev_max, "unit": "panel " + UNIT,
            "value_note": "panel haplotype-sites over the device time of the slowest rank (the same clock as the one-GPU line of this "
                          "config); the wall time beside it adds the host side of the call (stream upload, allocation, record download)",
            "records": nrec,
            "exchanges": {"all_gather_query_aggregate_bytes_per_rank": int(qagg.numel()) * 4, "all_gather_panel_aggregate_bytes_per_rank": int(pagg.numel()) * 4,
                          "state_hand_off_bytes": 8 * Q},
            "parity_vs_one_gpu": parity, "one_gpu_kernel_ms_same_box": one_gpu_ms,
            "strong_scaling_efficiency_kernels": (one_gpu_ms / (dev_max * 1e3) / world) if one_gpu_ms else None,
            "roofline_per_gpu": roofline_of((BYTES_AD * M * N + 16.0 * nrec) / world, dev_max * 1e3, peak),
            "kernel_ms_rank0": {k: v[0] for k, v in kt.items() if v[0] > 0}}


if __name__ == "__main__":
    main()
