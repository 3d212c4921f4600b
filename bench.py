#!/usr/bin/env python
"""bench.py — variants genotyped per second through the `PanGenie -f` hot path on B200.

One "step" = one pass of the whole stage (PRIME the graph k-mers, count the reads, histogram peak, fill, emission,
forward-backward, finalize) over ONE synthetic sample of the named BASELINE.json configuration (synthdata/large.py).

  workload  --gpus 1: configs[2] (22 chromosomes, 1 M variants, 32 haplotypes, 30x = 38 GB of FASTQ), the largest configuration that
            fits one GPU;  --gpus 4 / 8: configs[3] (5 M variants, 64 haplotypes, 30x = 189 GB of FASTQ), the north-star
            configuration, which needs the HBM of at least four GPUs;  --gpus 2: configs[2] again (configs[3] does not fit two).
            `--workload` overrides.
  value     whole-job variants/s with every input already resident in HBM
  e2e       the same through the reference-facing C-ABI with PINNED HOST buffers: host->device copies of reads / segments /
            panel and device->host copies of the results are inside the timed region
  multi-GPU ONE sample sharded over the ranks ("scaling": "strong"): chromosomes LPT-assigned, reads cut into record ranges;
            every rank PRIMEs the same canonical table, counts its shard, ONE all-reduce of the count array, no other
            collective
  roofline  dominant kernel by measured device time (CUDA events on the launching streams, inside the library):
            algorithmic bytes / duration against the measured HBM peak of MEASURED_PEAKS.json
  cpu_baseline / --impl reference
            the reference's own hmm.cpp (oracle/_ref) for emission + HMM and the CPU restatement of the jellyfish path for
            counting, on this box's host cores, on a bounded COMPLETE sub-sample (the smallest chromosomes with all their
            reads), extrapolated to the whole sample as stated in `sample`
  parity    GPU results of this run against the oracle on a slice (see `parity` in the line)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "variants genotyped per second (end-to-end PanGenie -f stage)"
HMM_KW = dict(recombrate=1.26, effective_N=1e-5)   # defaults of PanGenie (src/pangenie-genotype.cpp:33,42)
REGULARIZATION = 0.01
CPU_NS_PER_STATE = 165e-9                          # reference hmm.cpp, per state and column


def default_workload(n_gpus: int) -> str:
    return "cfg4" if n_gpus >= 4 else "cfg3"


def fb_bytes_per_column(P: int, A: int = 2) -> float:
    """B_fb = 2*8*P^2 + 2*2*P + 2*8*A^2 + 8 + 8*A(A+1)/2."""
    return 2 * 8 * P * P + 2 * 2 * P + 2 * 8 * A * A + 8 + 8 * A * (A + 1) / 2


def make_config(spec, world: int) -> dict:
    """Identical in both arms (the driver compares them)."""
    return {"workload": spec.text, "k": spec.k, "paths": spec.n_haplotypes + 1, "recombrate": HMM_KW["recombrate"],
            "effective_N": HMM_KW["effective_N"], "regularization": REGULARIZATION, "count_only_graph": True,
            "variants": "90% SNPs, 8% indels (1-50 bp), 2% tri-allelic, 1% with an undefined allele",
            "sample": "one sample; with several GPUs it is sharded over them (chromosomes LPT, reads by record ranges)",
            "l2": "inputs (tens of GB per GPU) are far larger than the 126 MB L2: no flush between steps"}


class ClockSampler:
    """Samples SM clock / throttle reasons DURING the timed region through NVML in-process, every 200 ms."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, device: int):
        self.device, self.sm, self.reasons, self.max_sm = device, [], set(), None
        self.period = float(os.environ.get("PG_BENCH_SAMPLE_PERIOD", "0.2"))
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
            for _ in range(3):  # the first NVML queries of a process are slow: keep them out of the timed region
                pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                reasons_fn(h)

            def loop():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        r = reasons_fn(h)
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
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload: str, kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture, or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        e = json.load(open(p)).get(workload, {}).get(kernel)
        return (e["bytes_per_launch"], e["source"]) if e else (None, None)
    except Exception:
        return None, None


def bind_numa(local: int) -> dict:
    """Pins this rank to the cores of its GPU's NUMA node BEFORE the pinned buffers are allocated (first touch), so the
    host->device streams of the ranks do not all cross one socket."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bdf = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return {"node": None}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus)}
    except Exception as e:
        return {"node": None, "error": type(e).__name__}


def cpu_model() -> str:
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return ""


# ------------------------------------------------------------------------------------------------------
# CPU arm: reference hmm.cpp (oracle/_ref) + CPU restatement of the jellyfish path.  Loads nothing of the product.
# ------------------------------------------------------------------------------------------------------
def panel_slice(p, n: int):
    """The first n variants of a panel (host arrays are views)."""
    from pangenie_b200.panel import Panel
    n = min(n, p.n_variants)
    K, A = int(p.kmer_offsets[n]), int(p.allele_offsets[n])
    F = int(p.flank_offsets[n]) if p.flank_offsets is not None else 0
    return Panel(p.n_paths, p.positions[:n], p.path_to_allele[:n * p.n_paths], p.coverage[:n], p.kmer_offsets[:n + 1], p.kmer_counts[:K],
                 p.allele_offsets[:n + 1], p.allele_ids[:A], p.allele_undefined[:A], p.allele_kmer_offset[:A], p.allele_kmer_mask[:A],
                 None if p.kmer_codes is None else p.kmer_codes[:K], None if p.flank_offsets is None else p.flank_offsets[:n + 1],
                 None if p.flank_codes is None else p.flank_codes[:F])


def choose_sample_chroms(spec, chrom_reads, record_bytes: int, budget_bytes: float = 1.5e9) -> list[int]:
    """The smallest chromosomes, as many as fit the read-byte budget (at least one)."""
    order = sorted(range(spec.n_chrom), key=lambda c: (chrom_reads[c], c))
    out, tot = [], 0
    for c in order:
        b = chrom_reads[c] * record_bytes
        if out and tot + b > budget_bytes:
            break
        out.append(c)
        tot += b
    return sorted(out)


def makespan(times, threads: int) -> float:
    """Wall time of a pool of `threads` workers taking the jobs in order (one job per chromosome, commands.cpp:949-978)."""
    free = [0.0] * max(1, threads)
    for t in times:
        i = min(range(len(free)), key=lambda j: free[j])
        free[i] += t
    return max(free)


def cpu_reference(spec, sample: dict, threads: int) -> dict:
    """One CPU step on the sub-sample + extrapolation to the whole sample.  `sample`: chroms, panels (one per sample
    chromosome, k-mer codes included), segments / reads (numpy u8 of exactly these chromosomes), chrom_variants (all),
    seg_bytes_total, read_bytes_total."""
    import ctypes as C
    from pangenie_b200.capi import PG_OP_PRIME, PG_OP_UPDATE, PgHmmParams, PgHmmResult, PgPanel, PgProbTable
    from pangenie_b200.panel import Result
    from tests import oracles
    oracle, ref = oracles.load_oracle(), oracles.load_ref()
    P = spec.n_haplotypes + 1
    segs, reads, panels = sample["segments"], sample["reads"], sample["panels"]
    complete = len(sample["chroms"]) == spec.n_chrom
    t0 = time.perf_counter()
    oc = oracles.OracleCounter(oracle, None, None, spec.k)
    oc.feed(segs, PG_OP_PRIME, threads=threads)
    t_prime = time.perf_counter() - t0
    t0 = time.perf_counter()
    oc.feed(reads, PG_OP_UPDATE, threads=threads)
    t_update = time.perf_counter() - t0
    t0 = time.perf_counter()
    peak = oc.computeHistogram(10000, True)
    t_hist = time.perf_counter() - t0
    t0 = time.perf_counter()
    oc.fill_counts(peak, panels)
    t_fill = time.perf_counter() - t0
    del oc
    # HMM: the reference's own class, one job per chromosome; long chromosomes are cut so the sample stays bounded
    cap_cols = max(500, int(float(os.environ.get("PG_BENCH_CPU_HMM_S", "12")) / (P * P * CPU_NS_PER_STATE)))
    hp = [panel_slice(p, cap_cols) for p in panels]
    table = PgProbTable()
    table.cov_min, table.cov_max, table.count_max, table.regularization = peak // 4, peak * 4, 2 * peak, REGULARIZATION
    prm = PgHmmParams()
    prm.recombrate, prm.effective_N, prm.uniform, prm.normalize = HMM_KW["recombrate"], HMM_KW["effective_N"], 0, 1
    res = [Result(p) for p in hp]
    pa, ra = (PgPanel * len(hp))(), (PgHmmResult * len(hp))()
    for i, (p, r) in enumerate(zip(hp, res)):
        pa[i], ra[i] = p.as_struct(), r.as_struct()
    secs = (C.c_double * len(hp))()
    hthreads = min(threads, spec.n_chrom)
    t0 = time.perf_counter()
    if ref is not None:
        f = ref.pgr_hmm_run_timed
        f.restype = C.c_int
        f.argtypes = [C.c_uint32, C.POINTER(PgPanel), C.POINTER(PgProbTable), C.POINTER(PgHmmParams), C.POINTER(PgHmmResult), C.c_int, C.POINTER(C.c_double)]
        st = f(len(hp), pa, C.byref(table), C.byref(prm), ra, min(hthreads, len(hp)), secs)
        kind = "reference"
    else:
        st = oracle.pgo_hmm_run_mt(len(hp), pa, C.byref(table), C.byref(prm), ra, min(hthreads, len(hp)))
        kind = "port"
    t_hmm_wall = time.perf_counter() - t0
    if st != 0:
        raise RuntimeError("CPU HMM failed")
    cols = sum(int(r.is_column.sum()) for r in res)
    hmm_cpu_s = sum(secs) if ref is not None and sum(secs) > 0 else t_hmm_wall * min(hthreads, len(hp))
    t_col = hmm_cpu_s / max(cols, 1)
    V_all = sample["chrom_variants"]
    fs, fr = sample["seg_bytes_total"] / max(len(segs), 1), sample["read_bytes_total"] / max(len(reads), 1)
    v_sample = sum(p.n_variants for p in panels)
    if complete and all(h.n_variants == p.n_variants for h, p in zip(hp, panels)):
        hmm_full, fill_full, extrap = t_hmm_wall, t_fill, "nothing extrapolated (the whole sample was processed)"
    else:
        # every variant with a non-reference allele on some path is an HMM column: columns scale like variants
        col_per_var = cols / max(sum(h.n_variants for h in hp), 1)
        hmm_full = makespan([v * col_per_var * t_col for v in V_all], hthreads)
        fill_full = makespan([v * t_fill / max(v_sample, 1) for v in V_all], hthreads)
        extrap = (f"extrapolated: PRIME x{fs:.1f} and histogram x{fs:.1f} (segment bytes), UPDATE x{fr:.1f} (read bytes), fill and HMM: measured time per "
                  f"variant / per column ({1e6 * t_col:.1f} us) applied to all {spec.n_chrom} chromosomes on a pool of {hthreads} threads")
    total = t_prime * fs + t_update * fr + t_hist * fs + fill_full + hmm_full
    V = int(sum(V_all))
    return {
        "value": V / total, "unit": "variants/s", "cores": threads, "cpu_model": cpu_model(), "kind": kind,
        "sample": (f"complete sub-sample = chromosome(s) {[c + 1 for c in sample['chroms']]} with all their reads ({len(reads) / 1e6:.0f} MB of "
                   f"{sample['read_bytes_total'] / 1e9:.1f} GB FASTQ, {len(segs) / 1e6:.0f} MB of {sample['seg_bytes_total'] / 1e6:.0f} MB segments, {v_sample} variants; HMM on "
                   f"the first {[h.n_variants for h in hp]} of them = {cols} columns); counting = CPU restatement of the jellyfish path (not libjellyfish) on {threads} "
                   f"threads, emission+HMM = reference hmm.cpp, one thread per chromosome; {extrap}"),
        "seconds_measured": {"prime": t_prime, "update": t_update, "histogram": t_hist, "fill": t_fill, "hmm_wall": t_hmm_wall, "hmm_cpu": hmm_cpu_s},
        "seconds_whole_sample": {"prime": t_prime * fs, "update": t_update * fr, "histogram": t_hist * fs, "fill": fill_full, "hmm": hmm_full, "total": total},
        "kmer_abundance_peak": int(peak),
        # which code each part of this baseline is: the whole-stage value mixes the reference's own emission + HMM with a
        # PORT of the jellyfish counting path; the reference-only part is reported on its own as well
        "parts": {"counting_fill_histogram": "port (oracle/pg_oracle.cpp: CPU restatement of the jellyfish 2.x path; libjellyfish is not in this image)",
                  "emission_hmm": "reference (unmodified src/hmm.cpp + emission / transition / indexer classes, oracle/_ref)" if ref is not None else "port"},
        "emission_hmm_only": {"value": V / hmm_full if hmm_full > 0 else None, "unit": "variants/s",
                              "what": "reference emission + forward-backward alone (makespan of its one-thread-per-chromosome pool); compare with "
                                      "variants / (emission_ms + hmm_skeleton_ms + hmm_blocks_ms + finalize_ms) of the GPU line"},
    }


def generate_cpu_sample(spec, device):
    """The sub-sample of the CPU arm, generated from scratch (reference arm process)."""
    from synthdata import large
    _lens, n_reads = large.chrom_plan(spec, device)
    RB = large.record_bytes(spec)
    chroms = choose_sample_chroms(spec, n_reads, RB)
    first = np.concatenate([[0], np.cumsum(n_reads)])
    segs, reads, panels = [], [], []
    for c in chroms:  # the chosen chromosomes need not be neighbours in the file: one generator call each
        w = large.make_workload(spec, device, chroms=[c], read_records=(int(first[c]), int(first[c + 1])), segment_chroms=[c])
        segs.append(w.segments.cpu().numpy())
        reads.append(w.reads.cpu().numpy())
        panels.append(w.panels[0])
    return {"chroms": chroms, "panels": panels, "segments": np.concatenate(segs), "reads": np.concatenate(reads),
            "chrom_variants": large.variants_per_chrom(spec), "seg_bytes_total": large.segments_bytes_estimate(spec, device),
            "read_bytes_total": int(sum(n_reads)) * RB}


def sample_from_workload(spec, wl):
    """The sub-sample of the CPU arm cut out of a fully generated workload (cpu_baseline of the B200 arm, N = 1)."""
    import copy
    RB = wl.record_bytes
    chroms = choose_sample_chroms(spec, wl.chrom_reads, RB)
    first = np.concatenate([[0], np.cumsum(wl.chrom_reads)])
    so = wl.segment_offsets
    segs = np.concatenate([wl.segments[so[c]:so[c + 1]].cpu().numpy() for c in chroms])
    reads = np.concatenate([wl.reads[int(first[c]) * RB:int(first[c + 1]) * RB].cpu().numpy() for c in chroms])
    panels = [copy.deepcopy(wl.panels[wl.my_chroms.index(c)]) for c in chroms]
    return {"chroms": chroms, "panels": panels, "segments": segs, "reads": reads, "chrom_variants": wl.chrom_variants,
            "seg_bytes_total": int(wl.segments.numel()), "read_bytes_total": int(sum(wl.chrom_reads)) * RB}


def run_reference(args, spec, config, W, K):
    import torch
    use_gpu = torch.cuda.is_available() and not os.environ.get("PG_BENCH_REF_GEN_CPU")
    device = torch.device("cuda", 0) if use_gpu else torch.device("cpu")
    t_all0 = time.perf_counter()
    sample = generate_cpu_sample(spec, device)   # data generation only; nothing below touches the GPU
    gen_s = time.perf_counter() - t_all0
    threads = os.cpu_count() or 1
    vals, last, executed = [], None, 0
    budget_s = float(os.environ.get("PG_BENCH_REF_BUDGET_S", "150"))
    t_run0 = time.perf_counter()
    for i in range(W + K):
        if vals and time.perf_counter() - t_run0 > budget_s:
            break
        last = cpu_reference(spec, sample, threads)
        if i >= W or time.perf_counter() - t_run0 > budget_s:
            vals.append(last["value"])
            executed += 1
    v = float(np.mean(vals))
    V = spec.n_variants
    line = {"metric": METRIC, "value": v, "unit": "variants/s", "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 * V / v,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f80 (x87 long double)", "data": "synthetic",
            "config": config, "impl": "reference", "cpu_baseline": {**last, "value": v},
            "e2e": {"value": v, "unit": "variants/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "steps_executed": executed, "wall_s": time.perf_counter() - t_all0, "data_generation_s": gen_s,
            "data_generation_device": str(device)}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# parity of this run against the oracle (outside the timed regions)
# ------------------------------------------------------------------------------------------------------
def parity_check(spec, wl, panels, peak, local):
    import copy
    import ctypes as C
    import pangenie_b200 as pg
    from pangenie_b200.capi import PG_OP_PRIME, PG_OP_UPDATE, PgHmmParams, PgHmmResult, PgPanel, PgProbTable
    from pangenie_b200.panel import Result
    from tests import oracles
    oracle, ref = oracles.load_oracle(), oracles.load_ref()
    out = {"oracle": ("reference hmm.cpp + ColumnIndexer + EmissionProbabilityComputer (oracle/_ref)" if ref is not None else "CPU restatement (oracle/pg_oracle.cpp)")
           + " for emission/HMM; CPU restatement of the jellyfish path for counting"}
    P = spec.n_haplotypes + 1
    # (1) emission + forward-backward: a slice of the smallest own chromosome with the counts this run filled
    ci = min(range(len(panels)), key=lambda i: panels[i].n_variants)
    n = min(panels[ci].n_variants, max(1500, int(4.0 / (P * P * CPU_NS_PER_STATE))))
    sl = copy.deepcopy(panel_slice(panels[ci], n))
    table = pg.ProbabilityTable(peak // 4, peak * 4, 2 * peak, REGULARIZATION)
    eng2 = pg.Engine(local)
    got = eng2.hmm_run([sl], table, **HMM_KW)[0]
    eng2.close()
    want = Result(sl)
    pa, ra = (PgPanel * 1)(), (PgHmmResult * 1)()
    pa[0], ra[0] = sl.as_struct(), want.as_struct()
    prm = PgHmmParams()
    prm.recombrate, prm.effective_N, prm.uniform, prm.normalize = HMM_KW["recombrate"], HMM_KW["effective_N"], 0, 1
    t = PgProbTable()
    t.cov_min, t.cov_max, t.count_max, t.regularization = peak // 4, peak * 4, 2 * peak, REGULARIZATION
    st = (ref.pgr_hmm_run if ref is not None else oracle.pgo_hmm_run)(1, pa, C.byref(t), C.byref(prm), ra)
    if st != 0:
        raise RuntimeError("oracle HMM failed")
    big = want.likelihoods > 1e-6
    rel = float(np.max(np.abs(got.likelihoods[big] - want.likelihoods[big]) / want.likelihoods[big])) if big.any() else 0.0
    out["hmm"] = {"chromosome": wl.my_chroms[ci] + 1, "variants": int(n), "columns": int(want.is_column.sum()),
                  "max_rel_err_likelihoods_above_1e-6": rel, "max_abs_err": float(np.max(np.abs(got.likelihoods - want.likelihoods))),
                  "gt_mismatches": int((got.genotype != want.genotype).reshape(-1, 2).any(axis=1).sum()),
                  "is_column_mismatches": int((got.is_column != want.is_column).sum()),
                  "gq_differs_by_more_than_1": int((np.abs(got.quality.astype(np.int64) - want.quality.astype(np.int64)) > 1)[want.quality < 150].sum())}
    # (2) counting + lookups + histogram: the first segment bytes of chromosome 1 and the first read bytes this rank holds
    so = wl.segment_offsets
    seg = wl.segments[so[0]:so[1]][:48 << 20].cpu().numpy()
    if len(seg) == 48 << 20:  # cut at a record boundary
        idx = np.flatnonzero((seg[1:] == ord(">")) & (seg[:-1] == 10))
        seg = seg[:int(idx[-1]) + 1]
    RB = wl.record_bytes
    rd = wl.reads[:((96 << 20) // RB) * RB].cpu().numpy()
    gc = pg.KmerCounter(rd, seg, spec.k, device=local)
    oc = oracles.OracleCounter(oracle, None, None, spec.k)
    thr = os.cpu_count() or 1
    oc.feed(seg, PG_OP_PRIME, threads=thr)
    oc.feed(rd, PG_OP_UPDATE, threads=thr)
    p0 = panels[0]
    codes = np.concatenate([p0.kmer_codes[:300_000], p0.flank_codes[:200_000]])
    a, b = gc.lookup(codes), oc.lookup(codes)
    hg, ho = gc.histogram(10000), oc.histogram(10000)
    out["counting"] = {"segment_bytes": int(len(seg)), "read_bytes": int(len(rd)), "kmers_compared": int(len(codes)),
                       "nonzero_counts": int((b > 0).sum()), "count_mismatches": int((a != b).sum()),
                       "histogram_equal": bool(np.array_equal(hg, ho)), "distinct_equal": bool(gc.distinct() == oc.distinct())}
    gc.close()
    out["ok"] = bool(out["hmm"]["gt_mismatches"] == 0 and out["hmm"]["is_column_mismatches"] == 0 and rel <= 1e-6
                     and out["counting"]["count_mismatches"] == 0 and out["counting"]["histogram_equal"])
    return out


# ------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------
def host_mem_available():
    """MemAvailable of /proc/meminfo in bytes (None if unreadable)."""
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                return int(ln.split()[1]) * 1024
    except OSError:
        pass
    return None


def run_b200(args, spec, name, config, rank, world, local, W, K):
    import torch
    import pangenie_b200 as pg
    from pangenie_b200.distributed import lpt_assign, sharded_count
    from pangenie_b200.panel import Result
    from synthdata import large
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    numa = bind_numa(local)
    lib = pg.load()
    dev = torch.device("cuda", local)
    per = large.variants_per_chrom(spec)
    mine = lpt_assign(per, world)[rank]
    _lens, n_reads = large.chrom_plan(spec, dev)
    total_reads = int(sum(n_reads))
    ra, rb = total_reads * rank // world, total_reads * (rank + 1) // world
    t0 = time.perf_counter()
    wl = large.make_workload(spec, dev, chroms=mine, read_records=(ra, rb))
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    torch.cuda.empty_cache()
    V, P = int(sum(per)), spec.n_haplotypes + 1
    panels = wl.panels
    reads_d, segs_d = wl.reads, wl.segments
    eng = pg.Engine(local)
    counter = None
    if world > 1:
        counter = pg.KmerCounter(None, None, spec.k, max_distinct=max(wl.segment_windows, 1024), device=local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    acc = {}

    def add(d):
        for k_, v_ in d.items():
            acc[k_] = acc.get(k_, 0) + v_

    def step(reads, segs, host: bool, results=None):
        """One whole `-f` stage.  host=True: pinned host text, panel upload and result download inside."""
        if world == 1:
            if host:
                _res, peak = eng.genotype_run(reads, segs, panels, k=spec.k, regularization=REGULARIZATION, results=results, **HMM_KW)
            else:
                peak = eng.run_resident(reads, segs, k=spec.k, regularization=REGULARIZATION, **HMM_KW)
            return peak, {}
        if host:
            eng.load(panels, results)
        counter.clear()
        t_c = time.perf_counter()
        info = sharded_count(counter, reads, segs, rank, world)
        info["count_wall_ms"] = 1e3 * (time.perf_counter() - t_c)
        probe_ms, probe_n = counter.last_probe_ms()
        info.update(probe_ms=probe_ms, probe_passes=probe_n, kmers=counter.kmers_seen())
        peak = eng.run_counted(counter, True, REGULARIZATION, **HMM_KW)
        if host:
            eng.fetch()
        return peak, info

    # ---------------- value: inputs resident in HBM ----------------
    eng.load(panels)
    for _ in range(W):
        step(reads_d, segs_d, False)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = lib.pg_kernel_launches()
    t0 = time.perf_counter()
    for _ in range(K):
        peak, info = step(reads_d, segs_d, False)
        add(eng.timings())
        add({"x_" + k_: v_ for k_, v_ in info.items()})
    barrier()
    dt = time.perf_counter() - t0
    launches = lib.pg_kernel_launches() - l0
    clocks = sampler.stop()
    eng.fetch()   # counts + results of the last step (parity check)

    # ---------------- e2e: pinned host buffers, copies inside the timed region ----------------
    e2e, dte = {}, None
    # All ranks decide TOGETHER whether the host can page-lock the inputs (the ranks of a box share its memory): a rank that
    # failed alone would leave the others in the barriers of the region, and a host driven out of memory kills the whole job.
    need = int(reads_d.numel() + segs_d.numel()) + (2 << 30)
    avail = host_mem_available()
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    fits = torch.tensor([1.0 if avail is None or avail / max(local_world, 1) > 1.25 * need else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(fits, op=dist.ReduceOp.MIN)
    try:
        if float(fits[0]) < 1.0:
            raise RuntimeError(f"host memory: {0 if avail is None else avail / 1e9:.0f} GB available for {local_world} ranks, "
                               f"{need / 1e9:.0f} GB of pinned buffers needed per rank")
        alloc_err = None
        try:
            reads_h = torch.empty(reads_d.numel(), dtype=torch.uint8, pin_memory=True)
            reads_h.copy_(reads_d)
            segs_h = torch.empty(segs_d.numel(), dtype=torch.uint8, pin_memory=True)
            segs_h.copy_(segs_d)
            torch.cuda.synchronize()
        except (RuntimeError, MemoryError) as ex:
            alloc_err = ex
        got = torch.tensor([0.0 if alloc_err is not None else 1.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(got, op=dist.ReduceOp.MIN)   # nobody enters the barriers of the region unless everybody can
        if float(got[0]) < 1.0:
            raise RuntimeError(f"pinned host buffers could not be allocated on some rank ({alloc_err})")
        for p in panels:  # the index arrays are uploaded every step: page-lock the large ones like the read buffers
            for nm in ("path_to_allele", "kmer_codes", "flank_codes", "positions"):
                a = getattr(p, nm)
                if a is not None and a.nbytes >= (1 << 20):
                    torch.cuda.cudart().cudaHostRegister(a.ctypes.data, a.nbytes, 0)
        res_bufs = [Result(p) for p in panels]
        Ke = min(K, int(os.environ.get("PG_BENCH_E2E_STEPS", str(K))))
        if Ke <= 0:
            raise RuntimeError("end-to-end region disabled (PG_BENCH_E2E_STEPS=0)")
        for _ in range(min(W, 2)):
            step(reads_h, segs_h, True, res_bufs)
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            peak_e, _i = step(reads_h, segs_h, True, res_bufs)
        barrier()
        dte = (time.perf_counter() - t0) / Ke
        assert peak_e == peak
        panel_names = ("positions", "path_to_allele", "kmer_offsets", "allele_offsets", "allele_ids", "allele_undefined", "allele_kmer_offset",
                       "allele_kmer_mask", "kmer_codes", "flank_offsets", "flank_codes")
        h2d = int(reads_h.numel() + segs_h.numel() + sum(sum(getattr(p, n_).nbytes for n_ in panel_names) for p in panels))
        d2h = int(sum(r.likelihoods.nbytes + r.is_column.nbytes + r.genotype.nbytes + r.quality.nbytes + r.unique_kmers.nbytes + r.coverage.nbytes
                      for r in res_bufs) + sum(p.kmer_counts.nbytes + p.coverage.nbytes for p in panels))
        e2e = {"h2d": h2d, "d2h": d2h, "steps": Ke}
        del reads_h
    except (RuntimeError, MemoryError) as ex:  # e.g. the host cannot page-lock tens of GB
        e2e = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}

    # ---------------- reduce over ranks ----------------
    t = torch.tensor([dt, dte if dte is not None else -1.0], dtype=torch.float64, device=dev)
    cnt = torch.tensor([float(launches), float(e2e.get("h2d", 0)), float(e2e.get("d2h", 0)), float(reads_d.numel())], dtype=torch.float64, device=dev)
    stage_names = ["prime_ms", "count_ms", "count_probe_ms", "histogram_ms", "fill_ms", "emission_ms", "hmm_skeleton_ms", "hmm_blocks_ms",
                   "finalize_ms", "x_prime_ms", "x_update_ms", "x_probe_ms", "x_exchange_ms", "x_count_wall_ms"]
    stages = torch.tensor([acc.get(n_, 0.0) / K for n_ in stage_names], dtype=torch.float64, device=dev)
    e2e_min = t[1:2].clone()
    stage_max = stages.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_min, op=dist.ReduceOp.MIN)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        dist.all_reduce(stage_max, op=dist.ReduceOp.MAX)
    dt, dte = float(t[0]), float(t[1])
    if float(e2e_min[0]) < 0:
        dte = None
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---------------- roofline (rank 0's kernels; stage maxima over ranks beside them) ----------------
    peak_gbs, peak_src = measured_peak_gbs()
    per_step = {n_: acc.get(n_, 0.0) / K for n_ in acc}
    text = float(reads_d.numel())
    if world == 1:
        upd_ms, prime_ms = per_step["count_ms"], per_step["prime_ms"]
        probe_ms, probe_n = per_step.get("count_probe_ms", 0.0), per_step.get("count_probe_passes", 0)
        kmers = per_step["kmers_counted"]
    else:
        upd_ms, prime_ms = per_step["x_update_ms"], per_step["x_prime_ms"]
        probe_ms, probe_n = per_step.get("x_probe_ms", 0.0), per_step.get("x_probe_passes", 0)
        kmers = per_step["x_kmers"]
    cols = per_step["hmm_columns"]
    tile_name = "count_tile_kernel<UPDATE> (parse, canonical k-mers" + (", scatter to the partition buffers)" if probe_n else ", probes)")
    kern = {
        "probe_parts_kernel<UPDATE>": {"ms": probe_ms, "launches": max(probe_n, 1), "alg_bytes": 16.0 * kmers,
                                       "what": "k-mer table probes + count increments of one super-chunk: 16 B per k-mer (8 B key probe + count read-modify-write)"},
        tile_name: {"ms": upd_ms - probe_ms, "launches": max(1, int(np.ceil(text / (64 << 20)))), "alg_bytes": text + (0.0 if probe_n else 16.0 * kmers),
                    "what": "FASTQ text streamed once" + ("" if probe_n else " + 16 B per k-mer")},
        "block_kernel (forward-backward + posterior)": {"ms": per_step["hmm_blocks_ms"], "launches": 1, "alg_bytes": fb_bytes_per_column(P) * cols,
                                                        "what": f"{fb_bytes_per_column(P):.0f} B per HMM column (alpha stored once, read once)"},
        "skeleton_kernel (sequential checkpoints)": {"ms": per_step["hmm_skeleton_ms"], "launches": 1, "alg_bytes": 0.0, "what": "latency-bound chain, no HBM roofline"},
        "count_tile_kernel<PRIME>": {"ms": prime_ms, "launches": max(1, int(np.ceil(segs_d.numel() / (64 << 20)))),
                                     "alg_bytes": float(segs_d.numel()) + 16.0 * wl.segment_windows, "what": "segment text + 16 B per graph k-mer"},
    }
    for kk in kern.values():
        kk["gbs"] = kk["alg_bytes"] / (kk["ms"] * 1e-3) / 1e9 if kk["ms"] > 0 else 0.0
    # The unit of SURVEY.md 8(d) for counting is the k-mer COUNTED: text byte read + 16 B of table per k-mer.  With partitioned
    # counting that unit passes through two kernels (scatter, then probe), so when the UPDATE pass dominates the step the
    # roofline is stated for the pass (both kernels, all their launches), with the per-kernel split beside it.
    upd_alg = text + 16.0 * kmers
    fbk = kern["block_kernel (forward-backward + posterior)"]
    fb_stage_ms = per_step["hmm_blocks_ms"] + per_step["hmm_skeleton_ms"]
    n_sets = max(1, int(np.ceil(text / (64 << 20))))
    candidates = {"UPDATE pass (count_tile_kernel<UPDATE" + (",SCATTER> + probe_parts_kernel<UPDATE>)" if probe_n else ">)"):
                  {"ms": upd_ms, "alg_bytes": upd_alg, "launches": (n_sets + probe_n) if probe_n else n_sets,
                   "what": "text streamed once + 16 B per k-mer (8 B key probe + count read-modify-write), SURVEY.md 8(d)"},
                  "block_kernel (forward-backward + posterior)": {**fbk, "what": fbk["what"]},
                  "count_tile_kernel<PRIME>": kern["count_tile_kernel<PRIME>"]}
    for kk in candidates.values():
        kk["gbs"] = kk["alg_bytes"] / (kk["ms"] * 1e-3) / 1e9 if kk["ms"] > 0 else 0.0
    dom = max(candidates, key=lambda n_: candidates[n_]["ms"])
    d = candidates[dom]
    traffic = traffic_src = None
    if dom.startswith("UPDATE") and probe_n:
        t_probe, s_probe = ncu_traffic(name, "probe_parts_kernel<UPDATE>")
        t_scat, _s = ncu_traffic(name, "count_tile_kernel<UPDATE>")
        if t_probe and t_scat:   # per step: every probe pass + every scatter launch set
            traffic = t_probe * probe_n + t_scat * n_sets
            traffic_src = "profiles/r2_ncu_cfg3.md: ncu --set full DRAM bytes per launch x launches per step (probe passes + scatter launch sets)"
    roofline = {"kernel": dom, "bound": "hbm", "achieved": d["gbs"], "peak": peak_gbs, "unit": "GB/s", "frac": d["gbs"] / peak_gbs,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "launches_per_step": d["launches"],
                "algorithmic_bytes_per_step": d["alg_bytes"], "ms_per_step": d["ms"], "algorithmic_bytes": d["what"],
                "rank": 0,
                "update_pass": {"achieved": upd_alg / (upd_ms * 1e-3) / 1e9 if upd_ms > 0 else 0.0,
                                "frac": (upd_alg / (upd_ms * 1e-3) / 1e9 / peak_gbs) if upd_ms > 0 else 0.0,
                                "ms": upd_ms, "probe_ms": probe_ms, "probe_passes": probe_n, "kmers": kmers, "text_bytes": text,
                                "algorithmic_bytes": "text + 16 B per k-mer (SURVEY.md 8d), all kernels of the UPDATE pass"},
                "forward_backward": {"achieved": fbk["gbs"], "frac": fbk["gbs"] / peak_gbs, "bytes_per_column": fb_bytes_per_column(P), "columns": cols,
                                     "ms": per_step["hmm_blocks_ms"], "skeleton_ms": per_step["hmm_skeleton_ms"],
                                     "stage_frac": (fbk["alg_bytes"] / (fb_stage_ms * 1e-3) / 1e9 / peak_gbs) if fb_stage_ms > 0 else 0.0},
                "kernels": {n_: {"ms": k_["ms"], "GBps": k_["gbs"], "frac": k_["gbs"] / peak_gbs, "algorithmic_bytes": k_["what"]} for n_, k_ in kern.items()},
                "stage_ms_rank0": {n_: per_step.get(n_, 0.0) for n_ in stage_names},
                "stage_ms_max_over_ranks": {n_: float(v_) for n_, v_ in zip(stage_names, stage_max.tolist())}}

    parity = None
    if not args.no_parity:
        try:
            parity = parity_check(spec, wl, panels, int(peak), local)
        except Exception as ex:
            parity = {"ok": False, "error": f"{type(ex).__name__}: {str(ex)[:300]}"}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            cpu = cpu_reference(spec, sample_from_workload(spec, wl), os.cpu_count() or 1)
        except Exception as ex:
            cpu = {"error": f"{type(ex).__name__}: {str(ex)[:300]}"}

    line = {"metric": METRIC, "value": V * K / dt, "unit": "variants/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": 1e3 * dt / K,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "clocks": clocks,
            "e2e": ({"value": V / dte, "unit": "variants/s", "h2d_bytes_per_step": int(cnt[1].item()), "d2h_bytes_per_step": int(cnt[2].item()),
                     "ms_per_step": 1e3 * dte, "steps": e2e.get("steps")} if dte else {"value": None, "unit": "variants/s", "unavailable": e2e.get("error")}),
            "gpu_launches": int(cnt[0].item()), "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
            "kmer_abundance_peak": int(peak), "workload": name,
            "multi_gpu": (f"{world} ranks, one sample: chromosomes LPT-sharded {[len(x) for x in lpt_assign(per, world)]}, reads sharded by record ranges, every rank "
                          "PRIMEs the canonical table, ONE all-reduce of the count array (NCCL, in 1 GiB pieces), no other collective on the data path")
            if world > 1 else "single GPU",
            "read_bytes_total": int(cnt[3].item()), "segment_bytes": int(segs_d.numel()), "numa": numa, "data_generation_s": gen_s}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("PG_BENCH_WORKLOAD", "auto"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from synthdata import large
    name = default_workload(max(args.gpus, world)) if args.workload == "auto" else args.workload
    if name not in large.CONFIGS:
        raise SystemExit(f"unknown workload {name}; choose from {sorted(large.CONFIGS)}")
    spec = large.CONFIGS[name]
    config = make_config(spec, world)
    # timing rule: at least 3 warm-up steps (PG_BENCH_MIN_WARMUP exists for runs under a profiler, whose numbers are never reported)
    W = max(args.warmup, int(os.environ.get("PG_BENCH_MIN_WARMUP", "3"))) if args.impl == "b200" else args.warmup
    K = max(1, args.steps)
    if args.impl == "reference":
        if rank == 0:
            run_reference(args, spec, config, W, K)
        return
    run_b200(args, spec, name, config, rank, world, local, W, K)


if __name__ == "__main__":
    main()
