#!/usr/bin/env python
"""bench.py -- headline benchmark of the PBWT build + maxWithin hot path (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one pass of the hot path over one synthetic panel: build (a[], d[], pack3 stream) +
matchMaximalWithin over M x N haplotype-sites.  metric = haplotype-sites/s.
  value : panel resident in HBM, CUDA events on the launching stream around each step.
  e2e   : the same step through the public host-buffer API: pinned host bit columns -> H2D ->
          kernels -> D2H of pack3 stream, aFend, d[] and every match record.
  roofline / cpu_baseline: see DESIGN.md section "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "haplotype-sites/sec build+maxWithin"
UNIT = "haplotype-sites/s"
M_DEFAULT, N_DEFAULT = 10000, 100000  # BASELINE.json configs[1]
SEED, L_RECOMB, MU = 0x5EED0002, 2000, 2.0
BYTES_PER_HS = 16.125  # read+write a, read+write d, one allele bit


def workload_name(M, N):
    return f"synthetic mosaic panel {M} haplotypes x {N} biallelic sites, build + maxWithin"


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


def time_reference(xbits, M, N_cpu):
    """The reference's own single-threaded CPU path on the first N_cpu sites (oracle/_ref when it was
    compiled here and travelled, else the oracle port)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O

    ref = O.reference()
    kind = "reference" if ref is not None else "port"
    imp = ref if ref is not None else O.port()
    tb, tm, nyz, nmatch = imp.time_build_maxwithin(xbits[:N_cpu], M, N_cpu)
    return {"value": M * N_cpu / (tb + tm), "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"first {N_cpu} of the workload's sites, all {M} haplotypes: build {tb:.2f}s + maxWithin {tm:.2f}s, "
                      f"{nmatch} matches (counting callback), single thread",
            "seconds": tb + tm}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--haps", type=int, default=M_DEFAULT)
    ap.add_argument("--sites", type=int, default=N_DEFAULT)
    ap.add_argument("--cpu-sites", type=int, default=0, help="sites timed on the CPU baseline (0 = automatic)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    M, N = args.haps, args.sites
    warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        if rank != 0:
            return
        # bit-identical panel from the CPU generator (no GPU needed on this arm)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O

        n_cpu = args.cpu_sites or min(N, 10000)
        x = O.generate_panel(M, n_cpu, SEED, L_RECOMB, MU)
        times = []
        for i in range(args.warmup + args.steps):
            cb = time_reference(x, M, n_cpu)
            if i >= args.warmup:
                times.append(cb["seconds"])
        t = float(np.mean(times)) if times else cb["seconds"]
        val = M * n_cpu / t
        cb.pop("seconds")
        cb["value"] = val
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                          "config": {"workload": workload_name(M, N), "haplotypes": M, "sites": N,
                                     "step": f"bounded sample: first {n_cpu} sites",
                                     "generator": f"seed {SEED:#x}, L={L_RECOMB}, mu={MU}"},
                          "cpu_baseline": cb,
                          "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import pbwt_b200 as P

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

    # N GPUs: the SITES of one panel of world*N sites are partitioned across the ranks (weak scaling: N sites per
    # GPU); the only exchange is one NCCL all-gather of per-shard aggregates (include/pbwtgpu.h, multi-GPU)
    n_total = N * world
    panel = eng.panel_synthetic(M, N, seed=SEED, L=L_RECOMB, mu=MU, first_site=rank * N)
    eng.synchronize()
    eng.last_timings()
    flags = P.WANT_YZ | P.WANT_MAXWITHIN
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

    # ---- end to end through the host-buffer API ----------------------------------------------------
    x_host = torch.empty((N, P.words_per_col(M)), dtype=torch.int32).pin_memory()
    x_np = x_host.numpy().view(np.uint32)
    x_np[:] = panel.to_host()
    pin = {
        "yz": torch.empty(max(nyz, 1), dtype=torch.uint8).pin_memory().numpy(),
        "a_end": torch.empty(M, dtype=torch.int32).pin_memory().numpy(),
        "d_end": torch.empty(M + 1, dtype=torch.int32).pin_memory().numpy(),
        "matches": torch.empty((max(nmatch, 1), 4), dtype=torch.int32).pin_memory().numpy().view(P.MATCH_DTYPE).reshape(-1),
    }
    h2d = x_np.nbytes
    d2h = nyz + 4 * M + 4 * (M + 1) + 16 * nmatch

    def e2e_step():
        pnl = eng.panel_from_host(x_np, M)
        res = device_step(pnl)
        res.fetch(out=pin)
        res.free()
        pnl.free()

    e2e_step()
    barrier()
    t0 = time.time()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    t_e2e_serial = (time.time() - t0) / args.steps
    t_e2e, in_flight = t_e2e_serial, 1

    # Two builds in flight (single GPU): two contexts, each with its own stream and its own host thread, take the
    # steps alternately, so the PCIe copies of one step overlap the kernels of the other.  Every step still does its
    # own H2D, kernels and D2H inside the timed region.
    if world == 1:
        engs = [P.Engine(local_rank), P.Engine(local_rank)]
        pins = [pin, {k: torch.from_numpy(v.view(np.uint8)).clone().pin_memory().numpy().view(v.dtype).reshape(v.shape)
                      for k, v in pin.items()}]

        def worker(w, steps):
            e = engs[w]
            for _ in range(steps):
                pnl = e.panel_from_host(x_np, M)
                res = e.build_device(pnl, flags)
                res.fetch(out=pins[w])
                res.free()
                pnl.free()

        for w in (0, 1):  # warm both contexts (their device-memory caches, kernel attributes)
            worker(w, 1)
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
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        algo_bytes = BYTES_PER_HS * M * N + nyz + 16.0 * nmatch  # rank 0's shard; the roofline is per GPU
        # one unit of algorithmic work (a step) needs ALL sweep-kernel launches of that step (pass 1 + combine
        # rounds + pass 2), so the denominator is their summed duration per step, not one launch's
        sweep_ms_per_step = sweep_ms / args.steps
        achieved = algo_bytes / (sweep_ms_per_step * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    # dram__bytes_read+write of pass 2, the dominant launch (502.2 MB; ncu --set full capture in
                    # profiles/r01b_c2_pass2_ncu_full_summary.txt), plus the window sort (13.8 MB) and its transpose
                    # (read 125 MB + write 130 MB); only valid for this config
                    "traffic": 771.0e6 if (M, N) == (M_DEFAULT, N_DEFAULT) else None, "kernel": "sweep_smem_kernel",
                    "kernel_ms_per_step": sweep_ms_per_step, "kernel_launches_per_step": sweep_launches / args.steps,
                    "kernel_ms_per_step_by_role": {"pass1": kt["prepass"][0] / args.steps, "combine": kt["combine"][0] / args.steps,
                                                   "pass2": kt["sweep"][0] / args.steps},
                    "algorithmic_bytes_per_step": algo_bytes, "peak_source": peak_src,
                    "note": "algorithmic bytes (16.125 B per haplotype-site + emitted bytes) over the summed duration of the "
                            "launches that do a step's sweep work: pass 1 (window nodes by transpose + radix sort of window "
                            "keys), combine rounds and pass 2 (sweep_smem_kernel, the dominant launch); a[]/d[] stay resident "
                            "in shared memory, so DRAM traffic is far below the algorithmic bytes"}
        cpu = None
        if not args.no_cpu_baseline:
            n_cpu = args.cpu_sites or min(N, 20000)
            cpu = time_reference(x_np, M, n_cpu)
            cpu.pop("seconds")
        out = {
            "metric": METRIC, "value": total_hs / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_name(M, n_total), "haplotypes": M, "sites": n_total, "sites_per_gpu": N,
                       "parallelism": (f"sites partitioned across {world} GPUs, one NCCL all-gather of per-shard aggregates "
                                       f"({agg_ints * 4} B per rank) per step") if world > 1 else "single GPU",
                       "l2": "256 MiB device memset between timed steps; per-step CUDA events on the launching stream",
                       "matches_per_haplotype": nmatch_all / M, "pack3_bytes": nyz_all, "generator": f"seed {SEED:#x}, L={L_RECOMB}, mu={MU}"},
            "e2e": {"value": total_hs / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h_all,
                    "ms_per_step": ms_e2e, "builds_in_flight": in_flight, "ms_per_step_one_in_flight": t_e2e_serial * 1e3,
                    "note": "host bit columns (pinned) -> pbwtgpu_panel_from_host -> pbwtgpu_build_device -> pbwtgpu_result_fetch of "
                            "pack3 stream, aFend, d and all match records, every step; with 2 builds in flight two contexts "
                            "(streams, host threads) alternate steps so one step's PCIe copies overlap the other's kernels"},
            "gpu_launches": launches_all,
            "kernel_ms": {k: v[0] for k, v in kt.items() if v[1]},
            "clocks": clocks.summary(),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "wall_s_timed_region": t_wall,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
