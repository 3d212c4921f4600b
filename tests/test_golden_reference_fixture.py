"""Golden fixture of the reference at the jellyfish boundary (reference tests/CommandsTest.cpp:17-95):
inputs  tests/data/index_path_segments.fasta (PRIME), region-reads.fa (FASTQ, UPDATE), index_chr1_kmers.tsv.gz
expect  kmer_abundance_peak == 18 (CommandsTest.cpp:57) and the per-variant counts / local coverages stored in
        tests/data/region_UniqueKmersList.cereal — produced by the real PanGenie with libjellyfish.
The files under tests/golden/counting/ are byte copies of those reference test-data files (data, not source);
pangenie_b200/refindex.py decodes them."""
import os

import numpy as np
import pytest

import pangenie_b200 as pg
from pangenie_b200 import refindex
from tests import oracles
from tests.helpers import assert_results_close

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "counting")


def _load():
    k, panels, _ = refindex.read_unique_kmers_map(os.path.join(G, "index_UniqueKmersMap.cereal"))
    panel = refindex.attach_kmers_tsv(panels["chr1"], os.path.join(G, "index_chr1_kmers.tsv.gz"))
    _, filled, _ = refindex.read_unique_kmers_map(os.path.join(G, "region_UniqueKmersList.cereal"))
    reads = np.fromfile(os.path.join(G, "region-reads.fa"), np.uint8)
    segs = np.fromfile(os.path.join(G, "index_path_segments.fasta"), np.uint8)
    return k, panel, filled["chr1"], reads, segs


def test_archive_decoding():
    k, panel, want, reads, segs = _load()
    assert k == 31 and panel.n_variants == 2 and panel.n_paths == 215
    assert reads[0] == ord("@") and segs[0] == ord(">")
    assert np.array_equal(panel.allele_ids, want.allele_ids) and np.array_equal(panel.allele_kmer_mask, want.allele_kmer_mask)
    assert np.array_equal(panel.path_to_allele, want.path_to_allele)


def test_counting_restatement_reproduces_jellyfish_counts(oracle):
    k, panel, want, reads, segs = _load()
    o = oracles.OracleCounter(oracle, reads, segs, k)
    # The reference test hard-codes kmer_abundance_peak = 18 (CommandsTest.cpp:57) for the HMM it compares with.
    # The 313-bp region yields a histogram of ~370 keys with three near-equal local maxima; the reference's own
    # Histogram code applied to the restated counts picks 35, so the peak itself is NOT pinned by this fixture
    # (it is decided by a handful of graph k-mers outside the 154 pinned ones) — the counts and coverages are.
    assert o.computeHistogram(10000, True) in (18, 35)
    o.fill_counts(18, [panel])
    assert np.array_equal(panel.kmer_counts, want.kmer_counts)
    assert np.array_equal(panel.coverage, want.coverage)


@pytest.mark.gpu
def test_gpu_reproduces_jellyfish_counts_and_reference_hmm(engine, oracle):
    k, panel, want, reads, segs = _load()
    g = pg.KmerCounter(reads, segs, k, hash_size=100000)
    o = oracles.OracleCounter(oracle, reads, segs, k)
    assert np.array_equal(g.histogram(), o.histogram()) and g.computeHistogram(10000, True) == o.computeHistogram(10000, True)
    engine.fill_counts(g, 18, [panel])   # peak as hard-coded by the reference test (CommandsTest.cpp:57)
    assert np.array_equal(panel.kmer_counts, want.kmer_counts) and np.array_equal(panel.coverage, want.coverage)
    res = engine.hmm_run([panel], pg.ProbabilityTable(18 // 4, 18 * 4, 2 * 18, 0.01), recombrate=1.26, effective_N=0.00001)
    # P = 215 paths, up to 45 alleles per variant: the reference's HMM on the same filled panel
    table = pg.ProbabilityTable(18 // 4, 18 * 4, 2 * 18, 0.01)
    ref = oracles.load_ref()
    lib, pre = (ref, "pgr_") if ref is not None else (oracle, "pgo_")
    exp = oracles.cpu_hmm_run(lib, pre, [want], table, recombrate=1.26, effective_N=0.00001)[0]
    assert_results_close(res[0], exp, label="CommandsTest fixture")


def test_oracle_hmm_on_fixture_matches_reference(oracle, ref):
    k, panel, want, reads, segs = _load()
    table = pg.ProbabilityTable(18 // 4, 18 * 4, 2 * 18, 0.01)
    a = oracles.cpu_hmm_run(oracle, "pgo_", [want], table, recombrate=1.26, effective_N=0.00001)[0]
    b = oracles.cpu_hmm_run(ref, "pgr_", [want], table, recombrate=1.26, effective_N=0.00001)[0]
    assert_results_close(a, b, rtol=1e-12, label="fixture")


@pytest.mark.gpu
def test_cpp_host_runs_the_stage_from_the_index_files(engine):
    """integration/genotype_from_index.cpp: a C++ host over the C-ABI only (native index reader -> pg_genotype_run).
    Its output equals what the Python mirror computes from the same index prefix and reads."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "integration", "genotype_from_index")
    if not os.path.exists(exe):
        pytest.skip("integration/genotype_from_index not built (make tools)")
    out = subprocess.run([exe, os.path.join(G, "index"), os.path.join(G, "region-reads.fa")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    rows = [l.split("\t") for l in out.stdout.splitlines() if not l.startswith("#")]
    ix = pg.Index(os.path.join(G, "index"))
    panel = ix.panel(0)
    reads = np.fromfile(os.path.join(G, "region-reads.fa"), np.uint8)
    segs = np.fromfile(ix.segments_path, np.uint8)
    res, peak = engine.genotype_run(reads, segs, [panel], k=ix.kmer_size, recombrate=1.26, effective_N=0.00001)
    assert f"peak: {peak}" in out.stderr
    assert len(rows) == panel.n_variants == 2
    for v, row in enumerate(rows):
        assert row[0] == "chr1" and int(row[1]) == int(panel.positions[v])
        gt = tuple(int(x) for x in row[2].split("/")) if row[2] != "./." else (-1, -1)
        assert gt == tuple(int(x) for x in res[0].genotype[2 * v:2 * v + 2])
        assert int(row[3]) == int(res[0].quality[v]) and int(row[4]) == int(res[0].unique_kmers[v]) and int(row[5]) == int(res[0].coverage[v])
        lik = np.array([float(x) for x in row[6].split(",")])
        want = res[0].likelihoods[int(res[0].gl_offsets[v]):int(res[0].gl_offsets[v + 1])]
        np.testing.assert_allclose(lik, want, rtol=1e-5, atol=1e-12)


@pytest.mark.gpu
def test_cpp_sharded_host_matches_the_single_gpu_host():
    """integration/genotype_sharded.cpp (C-ABI + NCCL: canonical PRIME on every GPU, read ranges, one all-reduce of the count
    array) prints exactly what the single-GPU host prints; with two visible GPUs the sample is really sharded."""
    import subprocess
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    one, many = os.path.join(root, "integration", "genotype_from_index"), os.path.join(root, "integration", "genotype_sharded")
    if not (os.path.exists(one) and os.path.exists(many)):
        pytest.skip("integration tools not built (make tools)")
    args = [os.path.join(G, "index"), os.path.join(G, "region-reads.fa")]
    want = subprocess.run([one] + args, capture_output=True, text=True)
    assert want.returncode == 0, want.stderr
    for n in sorted({1, min(2, torch.cuda.device_count())}):
        got = subprocess.run([many] + args + [str(n)], capture_output=True, text=True, timeout=600)
        assert got.returncode == 0, got.stderr[-2000:]
        lines = [ln for ln in got.stdout.splitlines(keepends=True) if not ln.startswith("NCCL version")]   # NCCL_DEBUG=VERSION banner
        assert "".join(lines) == want.stdout, n
        assert got.stderr.strip().splitlines()[-1] == want.stderr.strip().splitlines()[-1]   # the k-mer abundance peak
