"""The `PanGenie -f` stage end to end through pg_genotype_run (host buffers) vs the oracle pipeline."""
import numpy as np
import pytest

import pangenie_b200 as pg
from pangenie_b200 import synth
from tests import oracles
from tests.helpers import assert_results_close

pytestmark = pytest.mark.gpu


def _oracle_pipeline(oracle, wl, regularization=0.01, **kw):
    o = oracles.OracleCounter(oracle, wl.reads_fastq, wl.segments_fasta, wl.k)
    peak = o.computeHistogram(10000, True)
    o.fill_counts(peak, wl.panels)
    table = pg.ProbabilityTable(peak // 4, peak * 4, 2 * peak, regularization)
    return oracles.cpu_hmm_run(oracle, "pgo_", wl.panels, table, **kw), peak


def test_genotype_run_matches_oracle_pipeline(engine, oracle, tmp_path):
    wl = synth.make_workload(n_chrom=3, n_variants=1200, n_haplotypes=8, coverage=12.0, seed=21)
    kw = dict(recombrate=1.26, effective_N=1e-5)
    got, peak = engine.genotype_run(wl.reads_fastq, wl.segments_fasta, wl.panels, k=wl.k, histogram_path=str(tmp_path / "h.histo"), **kw)
    counts_gpu = [p.kmer_counts.copy() for p in wl.panels]
    cov_gpu = [p.coverage.copy() for p in wl.panels]
    want, peak_o = _oracle_pipeline(oracle, wl, **kw)
    assert peak == peak_o
    for p, c, cv in zip(wl.panels, counts_gpu, cov_gpu):
        assert np.array_equal(p.kmer_counts, c) and np.array_equal(p.coverage, cv)   # fill is bit-exact
    for i, (g, w) in enumerate(zip(got, want)):
        assert_results_close(g, w, label=f"chromosome {i}")
    lines = open(tmp_path / "h.histo").read().splitlines()
    assert len(lines) == 10002 and lines[-1].startswith("parameters\t")            # reference file format
    t = engine.timings()
    assert t["kernel_launches"] > 0 and t["hmm_columns"] > 0
    # the simulated sample is mostly recovered
    ok = tot = 0
    for r, tr in zip(got, wl.truth):
        ok += int((r.genotype.reshape(-1, 2) == np.sort(tr.astype(np.int16), axis=1)).all(axis=1).sum()); tot += len(tr)
    assert ok / tot > 0.9


def test_fill_counts_standalone(engine, oracle):
    wl = synth.make_workload(n_chrom=2, n_variants=500, n_haplotypes=4, coverage=8.0, seed=22)
    g = pg.KmerCounter(wl.reads_fastq, wl.segments_fasta, wl.k)
    peak = g.computeHistogram(10000, True)
    engine.fill_counts(g, peak, wl.panels)
    got = [(p.kmer_counts.copy(), p.coverage.copy()) for p in wl.panels]
    o = oracles.OracleCounter(oracle, wl.reads_fastq, wl.segments_fasta, wl.k)
    assert o.computeHistogram(10000, True) == peak
    o.fill_counts(peak, wl.panels)
    for p, (c, cv) in zip(wl.panels, got):
        assert np.array_equal(p.kmer_counts, c) and np.array_equal(p.coverage, cv)


def test_resident_engine_reuses_the_panel_across_samples(engine, oracle):
    """pg_engine_load once, then several samples through pg_engine_run_resident: the cached column structure, the
    reused counter and the scan/sequential checkpoint paths must not leak state from one sample to the next."""
    import torch
    wl = synth.make_workload(n_chrom=2, n_variants=700, n_haplotypes=8, coverage=10.0, seed=31)
    kw = dict(recombrate=1.26, effective_N=1e-5)
    want, peak_o = _oracle_pipeline(oracle, wl, **kw)
    segs_d = torch.from_numpy(wl.segments_fasta).cuda()
    reads_d = torch.from_numpy(wl.reads_fastq).cuda()
    half = (len(wl.reads_fastq) // wl.record_bytes // 2) * wl.record_bytes
    half_d = torch.from_numpy(wl.reads_fastq[:half].copy()).cuda()
    engine.load(wl.panels)
    for reads, full in ((reads_d, True), (half_d, False), (reads_d, True)):
        peak = engine.run_resident(reads, segs_d, k=wl.k, **kw)
        got = engine.fetch()
        if full:
            assert peak == peak_o
            for i, (g, w) in enumerate(zip(got, want)):
                assert_results_close(g, w, label=f"chromosome {i}")
    # a different path subset on the same loaded panel invalidates the cached column structure
    sub = dict(kw, only_paths=[0, 2, 3, 5])
    want_sub, _ = _oracle_pipeline(oracle, wl, **sub)
    engine.run_resident(reads_d, segs_d, k=wl.k, **sub)
    for i, (g, w) in enumerate(zip(engine.fetch(), want_sub)):
        assert_results_close(g, w, label=f"subset, chromosome {i}")
