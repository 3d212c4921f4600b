#!/usr/bin/env python
"""bench.py — variants genotyped per second through the `PanGenie -f` hot path on B200.

One "step" = one pass of the whole stage (PRIME segments, count reads, histogram peak, fill, emission,
forward-backward, finalize) over one synthetic sample of the named workload.

  value   whole-job variants/s with every input already resident in HBM (pg_engine_run_resident)
  e2e     the same through the reference-facing C-ABI call pg_genotype_run with PINNED HOST buffers:
          host->device copies of reads / segments / panel and device->host copies of the results are
          inside the timed region
  roofline  dominant kernel (by measured device time): algorithmic bytes / CUDA-event duration vs the
          measured HBM peak of MEASURED_PEAKS.json
  cpu_baseline  the reference's own hmm.cpp (oracle/_ref) for emission+HMM and the CPU restatement of the
          jellyfish path for counting, timed on this box's host cores on a bounded sample

Multi-GPU (`torchrun ... bench.py --gpus N`): workloads with one chromosome cannot shard, so every rank
genotypes its own sample of the same shape (weak scaling: the production scenario of one index, many
samples, reference README.md:128); 22-chromosome workloads shard chromosomes LPT-wise and reads by record
ranges, with one NCCL broadcast (primed keys) and one all-reduce (counts) — see DESIGN.md.
"""
from __future__ import annotations

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

WORKLOADS = {
    # name: (n_chrom, n_variants, n_haplotypes, coverage)  — BASELINE.json configs[1..]
    "cfg2": (1, 10_000, 8, 10.0),
    "cfg2x8": (1, 80_000, 8, 10.0),
    "cfg3s": (22, 100_000, 32, 30.0),   # configs[2] at 1/10 of the variants (same per-column shape)
    "h64s": (22, 50_000, 64, 30.0),     # configs[3] shape at 1/100 of the variants
}
WORKLOAD_TEXT = {
    "cfg2": "synthetic 1 chrom, 10k variants, 8 haplotypes, 10x reads, k=31 (BASELINE.json configs[1])",
    "cfg2x8": "synthetic 1 chrom, 80k variants, 8 haplotypes, 10x reads, k=31",
    "cfg3s": "synthetic 22 chroms, 100k variants, 32 haplotypes, 30x reads, k=31 (configs[2] shape, 1/10 variants)",
    "h64s": "synthetic 22 chroms, 50k variants, 64 haplotypes, 30x reads, k=31 (configs[3] shape, 1/100 variants)",
}


# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from one `ncu --set full` capture, summed over its
# launches of ONE step (like `achieved`, which is per step): profiles/r1_ncu_cfg2.md (two UPDATE launch sets per step)
NCU_TRAFFIC = {("cfg2", "count_tile_kernel<UPDATE>"): (1.315594e9 + 0.509737e9) + (1.150345e9 + 0.442476e9)}


def fb_bytes_per_column(P: int, A: int = 2) -> float:
    """B_fb = 2*8*P^2 + 2*2*P + 2*8*A^2 + 8 + 8*A(A+1)/2."""
    return 2 * 8 * P * P + 2 * 2 * P + 2 * 8 * A * A + 8 + 8 * A * (A + 1) / 2


class ClockSampler:
    """Samples SM clock / throttle reasons DURING the timed region through NVML in-process (a `nvidia-smi -lms` child
    was observed to stall CUDA calls of the benchmarked process for ~100 ms at a time)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, device: int):
        self.device, self.sm, self.reasons, self.max_sm = device, [], set(), None
        self.period = float(os.environ.get("PG_BENCH_SAMPLE_PERIOD", "0.2"))
        self.query_ms = []
        self._stop = threading.Event()
        self.t = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = self.device
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.device])
                except Exception:
                    pass
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons_fn = pynvml.nvmlDeviceGetCurrentClocksEventReasons if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons") \
                else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            for _ in range(3):  # the first NVML queries of a process are slow (tens of ms): keep them out of the timed region
                pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                reasons_fn(h)

            def loop():
                while not self._stop.is_set():
                    try:
                        tq = time.perf_counter()
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        r = reasons_fn(h)
                        self.query_ms.append(1e3 * (time.perf_counter() - tq))
                        for bit, name in self.REASONS.items():
                            if r & bit:
                                self.reasons.add(name)
                    except Exception:
                        pass
                    self._stop.wait(self.period)
            self.t = threading.Thread(target=loop, daemon=True)
            self.t.start()
        except Exception as e:  # no NVML: report that instead of inventing numbers
            self.reasons.add(f"nvml unavailable: {type(e).__name__}")

    def stop(self) -> dict:
        self._stop.set()
        if self.t:
            self.t.join(timeout=1)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "nvml_query_ms_max": max(self.query_ms) if self.query_ms else None}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_workload(name: str, seed_offset: int = 0):
    from pangenie_b200 import synth
    n_chrom, n_var, n_hap, cov = WORKLOADS[name]
    return synth.make_workload(n_chrom=n_chrom, n_variants=n_var, n_haplotypes=n_hap, coverage=cov, seed=20260925 + 1 + seed_offset)


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own code (oracle/_ref) for emission + HMM, restated jellyfish path for counting
# ------------------------------------------------------------------------------------------------------
def cpu_pipeline(wl, threads: int, sample_bytes: int, filled_panels_ok: bool):
    from tests import oracles
    import pangenie_b200 as pg
    oracle = oracles.load_oracle()
    ref = oracles.load_ref()
    rec = wl.record_bytes
    total = len(wl.reads_fastq)
    sample = min(total, max(rec, (sample_bytes // rec) * rec))
    if total <= (512 << 20):
        sample = total  # small workload: count every read (a few seconds), so fill + HMM see the real counts
        filled_panels_ok = False
    t0 = time.perf_counter()
    oc = oracles.OracleCounter(oracle, None, None, wl.k)
    oc.feed(wl.segments_fasta, pg.PG_OP_PRIME, threads=threads)
    t_prime = time.perf_counter() - t0
    t0 = time.perf_counter()
    oc.feed(wl.reads_fastq[:sample], pg.PG_OP_UPDATE, threads=threads)
    t_update_sample = time.perf_counter() - t0
    t_update = t_update_sample * total / sample
    # the HMM needs counts of the FULL read set: taken from the panels as filled by the GPU run (bit-identical
    # to the oracle's, tests/test_gpu_pipeline.py) so the CPU arm genotypes the same filled panel
    t0 = time.perf_counter()
    synthetic_counts = False
    if not filled_panels_ok:
        peak = oc.computeHistogram(10000, True)
        oc.fill_counts(peak, wl.panels)
    t_fill = time.perf_counter() - t0
    if not filled_panels_ok and sample < total:
        # counts of a read SAMPLE are too low to be representative for the HMM (most lookups would leave the table):
        # time the HMM on Poisson counts of the workload's nominal coverage instead
        from pangenie_b200 import synth
        cov = WORKLOADS[[k for k, v in WORKLOADS.items() if v[1] == wl.n_variants][0]][3]
        synth.fill_synthetic_counts(np.random.default_rng(1), wl, peak=max(int(cov * 0.75), 4))
        synthetic_counts = True
    peak = max(int(np.median(np.concatenate([p.coverage for p in wl.panels]))), 4)
    table = pg.ProbabilityTable(peak // 4, peak * 4, 2 * peak, 0.01)
    hthreads = min(threads, len(wl.panels))
    t0 = time.perf_counter()
    if ref is not None:
        oracles.cpu_hmm_run(ref, "pgr_", wl.panels, table, threads=hthreads, recombrate=1.26, effective_N=1e-5)
        kind = "reference"
    else:
        oracles.cpu_hmm_run(oracle, "pgo_", wl.panels, table, threads=hthreads, recombrate=1.26, effective_N=1e-5)
        kind = "port"
    t_hmm = time.perf_counter() - t0
    t_total = t_prime + t_update + t_fill + t_hmm
    cpu_model = ""
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                cpu_model = ln.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return {
        "value": wl.n_variants / t_total, "unit": "variants/s", "cores": threads, "cpu_model": cpu_model, "kind": kind,
        "sample": (f"emission+HMM: reference hmm.cpp on the full panel, {hthreads} thread(s) (one per chromosome, commands.cpp:949-978), "
                   f"{t_hmm:.3f}s; counting: CPU restatement of the jellyfish path (not libjellyfish), {threads} threads, PRIME full segments "
                   f"{t_prime:.2f}s + UPDATE on the first {sample / 1e6:.1f} MB of {total / 1e6:.1f} MB reads {t_update_sample:.2f}s "
                   f"extrapolated linearly to {t_update:.2f}s; fill {t_fill:.3f}s"
                   + ("; HMM timed on synthetic Poisson counts (the sampled counts are not representative)" if synthetic_counts else "")),
        "seconds": {"prime": t_prime, "update_extrapolated": t_update, "fill": t_fill, "hmm": t_hmm},
    }


def run_strong(args, rank, world, local, W, K, config):
    """One sample sharded over `world` GPUs."""
    import torch
    import torch.distributed as dist
    import pangenie_b200 as pg
    from pangenie_b200.distributed import lpt_assign, record_ranges, sharded_count
    wl = load_workload(args.workload)  # the SAME sample on every rank
    V = wl.n_variants
    mine = lpt_assign([p.n_variants for p in wl.panels], world)[rank]
    panels = [wl.panels[i] for i in mine]
    a, b = record_ranges(wl.reads_fastq, world)[rank]
    reads_h = torch.from_numpy(wl.reads_fastq[a:b].copy()).pin_memory()
    segs_h = torch.from_numpy(wl.segments_fasta).pin_memory()
    eng = pg.Engine(local)
    kw = dict(recombrate=1.26, effective_N=1e-5)
    counter = pg.KmerCounter(None, None, wl.k, max_distinct=len(wl.segments_fasta), device=local)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def step(reads, segs, load_fetch):
        if load_fetch:
            eng.load(panels)
        counter.clear()
        sharded_count(counter, reads, segs, rank, world)
        peak = eng.run_counted(counter, True, 0.01, **kw) if panels else 0
        if load_fetch and panels:
            eng.fetch()
        return peak

    reads_d, segs_d = reads_h.cuda(), segs_h.cuda()
    eng.load(panels)
    for _ in range(W):
        step(reads_d, segs_d, False)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    launches = 0
    for _ in range(K):
        flush.zero_()
        step(reads_d, segs_d, False)
        launches += eng.timings()["kernel_launches"]
    barrier()
    dt = time.perf_counter() - t0
    clocks = sampler.stop()
    for _ in range(2):
        step(reads_h, segs_h, True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        flush.zero_()
        peak = step(reads_h, segs_h, True)
    barrier()
    dte = time.perf_counter() - t0
    tt = torch.tensor([dt, dte], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt, dte = float(tt[0]), float(tt[1])
    nb = torch.tensor([float(reads_h.numel() + (segs_h.numel() if rank == 0 else 0)), float(launches)], dtype=torch.float64, device="cuda")
    dist.all_reduce(nb, op=dist.ReduceOp.SUM)
    if rank == 0:
        line = {"metric": "variants genotyped per second (end-to-end PanGenie -f stage)", "value": V * K / dt, "unit": "variants/s",
                "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {**config, "multi_gpu": "one sample: chromosomes LPT-sharded, reads sharded by record ranges, NCCL broadcast of the primed "
                           "key array + all-reduce of the count array, no other collective", "l2_flush": "256 MiB memset between steps"},
                "clocks": clocks, "e2e": {"value": V * K / dte, "unit": "variants/s", "h2d_bytes_per_step": int(nb[0].item()), "d2h_bytes_per_step": None,
                                          "ms_per_step": 1e3 * dte / K},
                "gpu_launches": int(nb[1].item()), "roofline": None, "cpu_baseline": None, "kmer_abundance_peak": int(peak)}
        print(json.dumps(line))
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("PG_BENCH_WORKLOAD", "cfg2"), choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="auto", choices=["auto", "weak", "strong"],
                    help="strong: ONE sample sharded over the GPUs (LPT chromosomes, record-aligned read shards, NCCL key "
                         "broadcast + count all-reduce); weak: one sample per GPU.  auto = strong when the workload has at "
                         "least as many chromosomes as GPUs")
    ap.add_argument("--inflight", type=int, default=int(os.environ.get("PG_BENCH_INFLIGHT", "2")),
                    help="independent samples kept in flight per GPU in the timed regions (host threads x engines); 1 = one call at a time")
    ap.add_argument("--inflight-e2e", type=int, default=int(os.environ.get("PG_BENCH_INFLIGHT_E2E", "3")),
                    help="samples in flight in the end-to-end region (the PCIe transfer of one sample hides the stages of two others)")
    ap.add_argument("--cpu-sample-mb", type=float, default=24.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    K = args.steps
    n_chrom, n_var, n_hap, cov = WORKLOADS[args.workload]
    config = {"workload": WORKLOAD_TEXT[args.workload], "k": 31, "paths": n_hap + 1, "recombrate": 1.26, "effective_N": 1e-5,
              "regularization": 0.01, "count_only_graph": True,
              "multi_gpu": "one sample per GPU (weak scaling over samples), no collective on the data path"}

    # ------------------------------------------------------------------ reference arm (CPU only)
    if args.impl == "reference":
        if rank != 0:
            return
        wl = load_workload(args.workload)
        threads = os.cpu_count() or 1
        vals = []
        t_all0 = time.perf_counter()
        last = None
        # a CPU step takes about a second on cfg2: the arm stops after ~150 s of work (at least one warm-up-free timed step),
        # so that a large --steps does not turn it into a quarter of an hour; `steps_executed` says how many were run
        budget_s = float(os.environ.get("PG_BENCH_REF_BUDGET_S", "150"))
        executed = 0
        for i in range(W + K):
            if vals and time.perf_counter() - t_all0 > budget_s:
                break
            last = cpu_pipeline(wl, threads, int(args.cpu_sample_mb * 1e6), filled_panels_ok=False)
            if i >= W or time.perf_counter() - t_all0 > budget_s:
                vals.append(last["value"])
                executed += 1
        v = float(np.mean(vals)) if vals else float("nan")
        line = {"metric": "variants genotyped per second (end-to-end PanGenie -f stage)", "value": v, "unit": "variants/s",
                "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 * wl.n_variants / v, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f80 (x87 long double)", "data": "synthetic", "config": config,
                "impl": "reference", "cpu_baseline": {**last, "value": v},
                "e2e": {"value": v, "unit": "variants/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "steps_executed": executed, "wall_s": time.perf_counter() - t_all0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import pangenie_b200 as pg
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    strong = world > 1 and (args.scaling == "strong" or (args.scaling == "auto" and n_chrom >= world))
    if strong:
        run_strong(args, rank, world, local, W, K, config)
        return
    wl = load_workload(args.workload, seed_offset=rank)  # every rank its own sample of the same shape
    V = wl.n_variants
    eng = pg.Engine(local)
    kw = dict(recombrate=1.26, effective_N=1e-5)
    reads_h = torch.from_numpy(wl.reads_fastq).pin_memory()
    segs_h = torch.from_numpy(wl.segments_fasta).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Samples are independent ("one index, thousands of samples", reference README.md:128): the timed regions keep
    # `--inflight` samples in flight, one host thread + one engine (own streams, own k-mer table) each, so the read transfer /
    # counting of one sample overlaps the forward-backward stage of another.  Every step still copies its own inputs and
    # results.  `serial` numbers (one sample at a time, the latency of one call) are reported beside them.
    import copy
    D = max(1, args.inflight)
    De = max(1, args.inflight_e2e)
    engines = [eng] + [pg.Engine(local) for _ in range(max(D, De) - 1)]
    panel_sets = [wl.panels] + [copy.deepcopy(wl.panels) for _ in range(max(D, De) - 1)]

    def run_in_flight(step_fn, steps, D=D):
        """steps calls of step_fn(lane) spread over D host threads; returns wall seconds (device-synchronised)."""
        per = [steps // D + (1 if i < steps % D else 0) for i in range(D)]
        errs = []

        def worker(i):
            try:
                torch.cuda.set_device(local)
                for _ in range(per[i]):
                    flush.zero_()  # evict L2 between steps (256 MiB > 126 MB L2)
                    step_fn(i)
            except Exception as ex:  # surface worker failures in the main thread
                errs.append(ex)
        barrier()
        t_start = time.perf_counter()
        if D == 1:
            worker(0)
        else:
            ths = [threading.Thread(target=worker, args=(i,)) for i in range(D)]
            for th in ths:
                th.start()
            for th in ths:
                th.join()
        barrier()
        if errs:
            raise errs[0]
        return time.perf_counter() - t_start

    # ---- value: inputs resident in HBM ----
    reads_d = reads_h.cuda()
    segs_d = segs_h.cuda()
    for e_, ps_ in zip(engines[:D], panel_sets[:D]):
        e_.load(ps_)
    for _ in range(W):
        for e_ in engines[:D]:
            e_.run_resident(reads_d, segs_d, k=wl.k, **kw)
    # a step lasts a few ms: keep warming up until the clocks have ramped (at least 0.3 s of work, still untimed)
    t_w = time.perf_counter()
    while time.perf_counter() - t_w < float(os.environ.get("PG_BENCH_WARM_S", "0.3")):
        eng.run_resident(reads_d, segs_d, k=wl.k, **kw)
    sampler = ClockSampler(local)
    sampler.start()
    dt = run_in_flight(lambda i: engines[i].run_resident(reads_d, segs_d, k=wl.k, **kw), K)
    clocks = sampler.stop()
    # serial pass without the NVML sampler thread: per-stage device times for the roofline (kernels of one sample only on the
    # GPU, so the CUDA-event times are clean) and the latency of one call
    tm_acc = {}
    barrier()
    t0u = time.perf_counter()
    for _ in range(K):
        flush.zero_()
        eng.run_resident(reads_d, segs_d, k=wl.k, **kw)
        t = eng.timings()
        for k_, v_ in t.items():
            tm_acc[k_] = tm_acc.get(k_, 0) + v_
    barrier()
    serial_ms = 1e3 * (time.perf_counter() - t0u) / K
    clocks["ms_per_step_without_sampler"] = serial_ms
    eng.fetch()
    tmax = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dt = float(tmax.item())
    value = world * V * K / dt

    # ---- e2e: pinned host buffers through pg_genotype_run, copies inside the timed region ----
    from pangenie_b200.panel import Result
    res_bufs = [[Result(p) for p in ps_] for ps_ in panel_sets]  # caller-owned output buffers, reused across steps
    out = [None] * len(engines)

    def e2e_step(i):
        out[i] = engines[i].genotype_run(reads_h, segs_h, panel_sets[i], k=wl.k, results=res_bufs[i], **kw)
    for _ in range(2):
        for i in range(De):
            e2e_step(i)
    dte = run_in_flight(e2e_step, K, De)
    res_e2e, peak = out[0]
    barrier()
    t0s = time.perf_counter()
    for _ in range(min(K, 20)):
        flush.zero_()
        e2e_step(0)
    barrier()
    e2e_serial_ms = 1e3 * (time.perf_counter() - t0s) / min(K, 20)
    tmax = torch.tensor([dte], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dte = float(tmax.item())
    e2e_value = world * V * K / dte
    panel_bytes = sum(sum(getattr(p, n).nbytes for n in ("positions", "path_to_allele", "kmer_offsets", "allele_offsets", "allele_ids",
                                                           "allele_undefined", "allele_kmer_offset", "allele_kmer_mask", "kmer_codes",
                                                           "flank_offsets", "flank_codes")) for p in wl.panels)
    h2d = int(reads_h.numel() + segs_h.numel() + panel_bytes)
    d2h = int(sum(r.likelihoods.nbytes + r.is_column.nbytes + r.genotype.nbytes + r.quality.nbytes + r.unique_kmers.nbytes + r.coverage.nbytes
                  for r in res_e2e) + sum(p.kmer_counts.nbytes + p.coverage.nbytes for p in wl.panels))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (per launch = per step, CUDA events on the launching stream) ----
    peak_gbs, peak_src = measured_peak_gbs()
    per = {k_: v_ / K for k_, v_ in tm_acc.items()}
    P = n_hap + 1
    kernels = {
        "count_tile_kernel<UPDATE>": {"ms": per["count_ms"], "alg_bytes": per["text_bytes"] + 16.0 * per["kmers_counted"]},
        "block_kernel (forward-backward)": {"ms": per["hmm_blocks_ms"], "alg_bytes": fb_bytes_per_column(P) * per["hmm_columns"]},
        "skeleton_kernel": {"ms": per["hmm_skeleton_ms"], "alg_bytes": 0.0},
        "count_tile_kernel<PRIME>": {"ms": per["prime_ms"], "alg_bytes": float(len(wl.segments_fasta)) * 17.0},
        "fill": {"ms": per["fill_ms"], "alg_bytes": 0.0}, "emission+descriptors": {"ms": per["emission_ms"], "alg_bytes": 0.0},
    }
    for kk in kernels.values():
        kk["gbs"] = kk["alg_bytes"] / (kk["ms"] * 1e-3) / 1e9 if kk["ms"] > 0 else 0.0
    dom = max(("count_tile_kernel<UPDATE>", "block_kernel (forward-backward)"), key=lambda n: kernels[n]["ms"])
    roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["gbs"], "peak": peak_gbs, "unit": "GB/s",
                "frac": kernels[dom]["gbs"] / peak_gbs, "traffic": NCU_TRAFFIC.get((args.workload, dom)),
                "traffic_source": "profiles/r1_ncu_cfg2.md (ncu --set full, per step)" if (args.workload, dom) in NCU_TRAFFIC else None,
                "launches_per_step": 2 if dom.startswith("count") and args.workload == "cfg2" else None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": kernels[dom]["alg_bytes"], "ms_per_launch": kernels[dom]["ms"],
                "forward_backward": {"achieved": kernels["block_kernel (forward-backward)"]["gbs"],
                                     "frac": kernels["block_kernel (forward-backward)"]["gbs"] / peak_gbs,
                                     "bytes_per_column": fb_bytes_per_column(P), "columns": per["hmm_columns"],
                                     "ms": per["hmm_blocks_ms"], "skeleton_ms": per["hmm_skeleton_ms"]},
                "stage_ms": {k_: per[k_] for k_ in ("prime_ms", "count_ms", "histogram_ms", "fill_ms", "emission_ms", "hmm_skeleton_ms",
                                                     "hmm_blocks_ms", "finalize_ms")}}

    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_pipeline(wl, os.cpu_count() or 1, int(args.cpu_sample_mb * 1e6), filled_panels_ok=True)

    line = {"metric": "variants genotyped per second (end-to-end PanGenie -f stage)", "value": value, "unit": "variants/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {**config, "l2_flush": "256 MiB memset between steps",
                       "samples_in_flight": f"{D} per GPU in the resident region, {De} in the end-to-end region (one host thread + one engine each); "
                                            "serial_ms_per_step = one call at a time"},
            "serial_ms_per_step": serial_ms,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "variants/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * dte / K,
                    "serial_ms_per_step": e2e_serial_ms},
            "gpu_launches": int(tm_acc["kernel_launches"]), "roofline": roofline, "cpu_baseline": cpu,
            "kmer_abundance_peak": int(peak)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
