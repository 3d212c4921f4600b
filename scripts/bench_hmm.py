#!/usr/bin/env python
"""Forward-backward only (pg_hmm_run on pre-filled panels): steady-state roofline of the block kernel at the
BASELINE.json haplotype counts, with enough columns that the persistent grid runs many waves of block jobs.

  python scripts/bench_hmm.py [--haplotypes 32 64 128] [--variants 400000]
Prints one JSON line per shape: device times from the library's CUDA events, algorithmic bytes of SURVEY.md 8(d)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--haplotypes", type=int, nargs="+", default=[32, 64, 128])
    ap.add_argument("--variants", type=int, default=400_000)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--samples", type=int, default=1,
                    help="samples genotyped in ONE pg_hmm_run call (SURVEY 8f row 4): the panels of every sample are passed as further "
                         "chromosomes, so all their forward / backward checkpoint walks run concurrently and the block kernel sees S times the jobs")
    args = ap.parse_args()
    import pangenie_b200 as pg
    from synthdata import small as synth
    from bench import fb_bytes_per_column, measured_peak_gbs
    peak_gbs, src = measured_peak_gbs()
    eng = pg.Engine(0)
    for H in args.haplotypes:
        V = args.variants if H <= 64 else args.variants // 4
        t0 = time.time()
        wl = synth.make_workload(n_chrom=22, n_variants=V, n_haplotypes=H, coverage=0, with_reads=False, seed=7 + H)
        synth.fill_synthetic_counts(np.random.default_rng(H), wl)
        panels = list(wl.panels)
        import copy
        for smp in range(1, args.samples):   # further samples on the same index: same structure, their own counts
            synth.fill_synthetic_counts(np.random.default_rng(1000 * smp + H), wl)
            panels += [copy.deepcopy(p_) for p_ in wl.panels]
        table = pg.ProbabilityTable(6, 96, 48, 0.01)
        kw = dict(recombrate=1.26, effective_N=1e-5)
        if H + 1 > 100:
            kw["only_paths"] = list(range(H + 1))  # "-a 129": one subset of all paths
        best = None
        for _ in range(args.repeat):
            eng.hmm_run(panels, table, **kw)
            t = eng.timings()
            if best is None or t["hmm_blocks_ms"] < best["hmm_blocks_ms"]:
                best = t
        cols = best["hmm_columns"]
        alg = fb_bytes_per_column(H + 1) * cols
        gbs = alg / (best["hmm_blocks_ms"] * 1e-3) / 1e9
        print(json.dumps({"haplotypes": H, "variants": V, "samples": args.samples, "columns": cols,
                          "stage_columns_per_s": cols / ((best["hmm_blocks_ms"] + best["hmm_skeleton_ms"]) * 1e-3),
                          "stage_frac": alg / ((best["hmm_blocks_ms"] + best["hmm_skeleton_ms"]) * 1e-3) / 1e9 / peak_gbs, "blocks_ms": best["hmm_blocks_ms"], "skeleton_ms": best["hmm_skeleton_ms"],
                          "emission_ms": best["emission_ms"], "algorithmic_GB": alg / 1e9, "achieved_GBps": gbs, "peak_GBps": peak_gbs,
                          "frac": gbs / peak_gbs, "peak_source": src, "columns_per_s_blocks": cols / (best["hmm_blocks_ms"] * 1e-3),
                          "setup_s": time.time() - t0}), flush=True)


if __name__ == "__main__":
    main()
