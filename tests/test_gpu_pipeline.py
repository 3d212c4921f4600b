"""The `PanGenie -f` stage end to end through pg_genotype_run (host buffers) vs the oracle pipeline."""
import numpy as np
import pytest

import pangenie_b200 as pg
from synthdata import small as synth
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


def test_configs1_full_size_properties(engine, monkeypatch):
    """BASELINE.json configs[1] at full size (1 chromosome, 10k variants, 8 haplotypes, 10x reads) through
    size-independent properties: the resident and the host-buffer paths agree bit for bit, the partitioned and the direct
    counting paths produce identical counts, rows are normalised, the run is reproducible, the sample is recovered."""
    import torch
    wl = synth.make_workload(n_chrom=1, n_variants=10_000, n_haplotypes=8, coverage=10.0, seed=20260926)
    kw = dict(recombrate=1.26, effective_N=1e-5)
    got, peak = engine.genotype_run(wl.reads_fastq, wl.segments_fasta, wl.panels, k=wl.k, **kw)
    lik = got[0].likelihoods.copy(); gt = got[0].genotype.copy(); gq = got[0].quality.copy()
    counts = wl.panels[0].kmer_counts.copy(); cov = wl.panels[0].coverage.copy()
    assert 5 <= peak <= 10                                           # ~7.5 expected
    sums = np.add.reduceat(lik, got[0].gl_offsets[:-1].astype(np.int64))
    assert np.allclose(sums[got[0].is_column == 1], 1.0, atol=1e-9)
    t = np.sort(wl.truth[0].astype(np.int16), axis=1)
    assert (gt.reshape(-1, 2) == t).all(axis=1).mean() > 0.9
    # resident path (device text), direct and partitioned counting
    reads_d = torch.from_numpy(wl.reads_fastq).cuda(); segs_d = torch.from_numpy(wl.segments_fasta).cuda()
    for part_kb in ("0", "16384"):
        monkeypatch.setenv("PG_COUNT_PART_KB", part_kb)
        engine.load(wl.panels)
        assert engine.run_resident(reads_d, segs_d, k=wl.k, **kw) == peak
        res = engine.fetch()[0]
        assert np.array_equal(wl.panels[0].kmer_counts, counts) and np.array_equal(wl.panels[0].coverage, cov)
        assert np.array_equal(res.likelihoods, lik) and np.array_equal(res.genotype, gt) and np.array_equal(res.quality, gq)
    assert engine.timings()["hmm_scan_used"] == 1
