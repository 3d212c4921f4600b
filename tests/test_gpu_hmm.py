"""GPU forward-backward / emission vs the CPU oracle (and the reference itself where oracle/_ref travelled).
Bar (BASELINE.json north_star): posteriors within 1e-6 relative, identical GT calls."""
import copy
import os
import subprocess

import numpy as np
import pytest

import pangenie_b200 as pg
from tests import oracles
from tests.helpers import assert_results_close, random_panel

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _table():
    return pg.ProbabilityTable(4, 72, 36, 0.01)


def test_reference_vector_get_genotyping_result(engine):
    # reference tests/HMMTest.cpp:14-46 ("computed by hand")
    b = pg.PanelBuilder()
    u1 = b.add_variant(2000, [0, 1]); b.insert_kmer(u1, 10, [0]); b.insert_kmer(u1, 10, [1]); b.set_coverage(u1, 5)
    u2 = b.add_variant(3000, [0, 1]); b.insert_kmer(u2, 20, [0]); b.insert_kmer(u2, 5, [1]); b.set_coverage(u2, 5)
    probs = pg.ProbabilityTable(5, 10, 30, 0.0)
    probs.modify_probability(5, 10, 0.1, 0.9, 0.1)
    probs.modify_probability(5, 20, 0.01, 0.01, 0.9)
    probs.modify_probability(5, 5, 0.9, 0.3, 0.1)
    res = engine.hmm_run([b.build()], probs, recombrate=446.287102628, effective_N=0.25)[0]
    expected = [0.0509465435, 0.9483202731, 0.0007331832, 0.9678020017, 0.031003181, 0.0011948172]
    got = [res.get_genotype_likelihood(v, a1, a2) for v in range(2) for (a1, a2) in ((0, 0), (0, 1), (1, 1))]
    assert np.abs(np.array(got) - expected).max() < 1e-7  # the reference's doubles_equal tolerance


def test_reference_vector_underflow_and_skip(engine):
    # reference tests/HMMTest.cpp:554-588 (underflow -> uniform) and :48-98 (skipped reference-only column)
    b = pg.PanelBuilder()
    for pos, c2 in ((1000, 10), (2000, 0), (3000, 10)):
        v = b.add_variant(pos, [0, 1]); b.insert_kmer(v, 10 if c2 else 20, [0]); b.insert_kmer(v, c2, [1])
    probs = pg.ProbabilityTable(0, 1, 21, 0.0)
    probs.modify_probability(0, 10, 0.0, 1.0, 0.0)
    probs.modify_probability(0, 20, 0.0, 0.0, 1.0)
    probs.modify_probability(0, 0, 1.0, 0.0, 0.0)
    res = engine.hmm_run([b.build()], probs, recombrate=0.0, effective_N=0.25)[0]
    got = [res.get_genotype_likelihood(v, a1, a2) for v in range(3) for (a1, a2) in ((0, 0), (0, 1), (1, 1))]
    assert np.abs(np.array(got) - [0, 0, 0, 0, 1, 0, 0, 1, 0]).max() < 1e-7


CASES = [
    dict(n_variants=50, n_paths=2),
    dict(n_variants=300, n_paths=9, max_alleles=2, shared_kmer_frac=0.3),          # cfg (4,3,1), H=8 shape
    dict(n_variants=200, n_paths=5, max_alleles=3, undefined_frac=0.2),
    dict(n_variants=150, n_paths=17, max_alleles=6, undefined_frac=0.1, kmers_per_allele=(0, 12)),  # general-A path
    dict(n_variants=300, n_paths=33, max_alleles=3),                               # cfg (4,9,1), H=32 shape
    dict(n_variants=200, n_paths=65, max_alleles=2),                               # cfg (4,17,1), H=64 shape
    dict(n_variants=70, n_paths=129, max_alleles=3),                               # cfg (8,17,2), H=128 shape
    dict(n_variants=12, n_paths=215, max_alleles=4),                               # spill config, P of the reference fixture
]


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("normalize", [True, False])
def test_hmm_matches_oracle(engine, oracle, case, normalize):
    rng = np.random.default_rng(1000 + case)
    panel = random_panel(rng, **CASES[case])
    table = _table()
    kw = dict(recombrate=1.26, effective_N=1e-5 if case % 2 == 0 else 25000.0, normalize=normalize)
    want = oracles.cpu_hmm_run(oracle, "pgo_", [panel], table, **kw)[0]
    got = engine.hmm_run([panel], table, **kw)[0]
    assert_results_close(got, want, rtol=1e-6 if normalize else 1e-6, atol=1e-300, label=f"case {case}")


def test_hmm_multi_chromosome_blocks_and_checkpoints(engine, oracle):
    # chains longer than several checkpoint blocks, several chromosomes at once
    rng = np.random.default_rng(5)
    panels = [random_panel(rng, n, 9, max_alleles=3, ref_only_frac=0.05) for n in (1500, 700, 1, 300)]
    table = _table()
    want = oracles.cpu_hmm_run(oracle, "pgo_", panels, table, recombrate=1.26, effective_N=1e-5)
    got = engine.hmm_run(panels, table, recombrate=1.26, effective_N=1e-5)
    for i, (g, w) in enumerate(zip(got, want)):
        assert_results_close(g, w, label=f"chromosome {i}")


@pytest.mark.parametrize("n_paths,block", [(9, 64), (9, 7), (4, 16), (2, 2), (7, 33)])
def test_scan_and_sequential_checkpoints_agree(engine, oracle, monkeypatch, n_paths, block):
    """P <= 9: checkpoints come from the parallel-in-time basis/scan kernels (csrc/hmm_scan.cuh); PG_SKELETON=seq forces
    the sequential skeleton walk.  Both must match the oracle, and each other far below the parity tolerance."""
    rng = np.random.default_rng(77 + n_paths + block)
    panels = [random_panel(rng, n, n_paths, max_alleles=4, undefined_frac=0.1, ref_only_frac=0.05) for n in (700, 65, 130)]
    table = _table()
    monkeypatch.setenv("PG_HMM_B", str(block))
    for normalize in (True, False):
        kw = dict(recombrate=1.26, effective_N=25000.0 if block % 2 else 1e-5, normalize=normalize)
        want = oracles.cpu_hmm_run(oracle, "pgo_", panels, table, **kw)
        monkeypatch.setenv("PG_SKELETON", "scan")
        got_scan = engine.hmm_run(panels, table, **kw)
        assert engine.timings()["hmm_scan_used"] == 1
        monkeypatch.setenv("PG_SKELETON", "seq")
        got_seq = engine.hmm_run(panels, table, **kw)
        assert engine.timings()["hmm_scan_used"] == 0
        for i, (a, b, w) in enumerate(zip(got_scan, got_seq, want)):
            assert_results_close(a, w, label=f"scan, chromosome {i}")
            assert_results_close(b, w, label=f"seq, chromosome {i}")
            np.testing.assert_allclose(a.likelihoods, b.likelihoods, rtol=1e-11, atol=1e-300)


def test_scan_falls_back_on_zero_totals(engine, oracle, monkeypatch):
    """Emission tables with exact zeros kill basis chains (and sometimes the real chain: the reference's uniform
    replacement, hmm.cpp:258-260) -> the flagged chromosome is recomputed sequentially; results stay exact."""
    rng = np.random.default_rng(123)
    probs = pg.ProbabilityTable(0, 1, 21, 0.0)
    probs.modify_probability(0, 10, 0.0, 1.0, 0.0)
    probs.modify_probability(0, 20, 0.0, 0.0, 1.0)
    probs.modify_probability(0, 0, 1.0, 0.0, 0.0)
    b = pg.PanelBuilder()
    pos = 1000
    for i in range(40):
        pos += int(rng.integers(100, 2000))
        al = rng.integers(0, 2, size=6)
        al[0], al[1] = 0, 1
        v = b.add_variant(pos, al)
        c = [(10, 10), (20, 0), (0, 20), (10, 0)][int(rng.integers(0, 4))]
        b.insert_kmer(v, c[0], [0]); b.insert_kmer(v, c[1], [1])
    panel = b.build()
    monkeypatch.setenv("PG_HMM_B", "4")
    monkeypatch.setenv("PG_SKELETON", "scan")
    kw = dict(recombrate=1.26, effective_N=25000.0)
    want = oracles.cpu_hmm_run(oracle, "pgo_", [panel], probs, **kw)[0]
    got = engine.hmm_run([panel], probs, **kw)[0]
    assert_results_close(got, want, label="zero-emission chain")


def test_hmm_options(engine, oracle):
    rng = np.random.default_rng(9)
    panel = random_panel(rng, 200, 8, max_alleles=3)
    table = _table()
    for kw in (dict(only_paths=[0, 3, 5, 7]), dict(uniform=True), dict(recombrate=446.287102628, effective_N=0.25),
               dict(only_paths=[1, 2], normalize=False)):
        want = oracles.cpu_hmm_run(oracle, "pgo_", [panel], table, **kw)[0]
        got = engine.hmm_run([panel], table, **kw)[0]
        assert_results_close(got, want, label=str(kw))


def test_path_subsets_are_combined_like_the_reference(engine, oracle):
    """`-a`: run_genotyping adds the un-normalised likelihoods of disjoint path subsets and normalises at the end
    (src/commands.cpp:155-176, 982-988).  Expected values: the oracle run per subset with normalize=False, summed."""
    rng = np.random.default_rng(17)
    panels = [random_panel(rng, n, 12, max_alleles=3, undefined_frac=0.05, ref_only_frac=0.05) for n in (260, 90)]
    table = _table()
    subsets = [[0, 4, 7, 9], [1, 2, 3, 11], [5, 6, 8, 10]]
    kw = dict(recombrate=1.26, effective_N=25000.0)
    got = engine.hmm_run_subsets(panels, table, subsets, **kw)
    for c, panel in enumerate(panels):
        total = None
        col_any = np.zeros(panel.n_variants, bool)
        for sub in subsets:
            r = oracles.cpu_hmm_run(oracle, "pgo_", [panel], table, only_paths=sub, normalize=False, **kw)[0]
            total = r.likelihoods.copy() if total is None else total + r.likelihoods
            col_any |= r.is_column == 1
        off = got[c].gl_offsets.astype(np.int64)
        sums = np.add.reduceat(total, off[:-1])
        norm = total / np.repeat(np.where(sums > 0, sums, 1.0), np.diff(off))
        assert np.array_equal(got[c].is_column == 1, col_any)
        np.testing.assert_allclose(got[c].likelihoods, norm, rtol=1e-6, atol=1e-300)
        # GT = the likeliest genotype of the combined likelihoods (biallelic and defined variants checked by argmax)
        for v in np.nonzero(col_any)[0][:200]:
            row = norm[off[v]:off[v + 1]]
            if len(row) == 3 and np.sort(row)[-1] - np.sort(row)[-2] > 1e-6 and not panel.allele_undefined[panel.allele_offsets[v]:panel.allele_offsets[v + 1]].any():
                g = [(0, 0), (0, 1), (1, 1)][int(np.argmax(row))]
                assert tuple(got[c].genotype[2 * v:2 * v + 2]) == g, v


def test_hmm_matches_reference_sources_directly(engine, ref):
    rng = np.random.default_rng(11)
    panel = random_panel(rng, 120, 12, max_alleles=4, undefined_frac=0.15)
    table = _table()
    want = oracles.cpu_hmm_run(ref, "pgr_", [panel], table, recombrate=1.26, effective_N=1e-5)[0]
    got = engine.hmm_run([panel], table, recombrate=1.26, effective_N=1e-5)[0]
    assert_results_close(got, want, label="vs reference hmm.cpp")


def test_emission_matches_oracle(engine, oracle):
    rng = np.random.default_rng(3)
    panel = random_panel(rng, 300, 7, max_alleles=4, undefined_frac=0.3, shared_kmer_frac=0.5, count_range=(0, 200))
    table = pg.ProbabilityTable(4, 30, 40, 0.01)
    o1, e1, l1 = engine.emission_run(panel, table)
    o2, e2, l2 = oracles.cpu_emission_run(oracle, "pgo_", panel, table)
    assert np.array_equal(o1, o2)
    np.testing.assert_allclose(e1, e2, rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(l1, l2, rtol=1e-10, atol=1e-10)


def test_empty_and_degenerate_inputs(engine, oracle):
    table = _table()
    b = pg.PanelBuilder()
    v = b.add_variant(100, [0, 0, 0])  # reference-only: no HMM column at all
    b.insert_kmer(v, 3, [0])
    res = engine.hmm_run([b.build()], table)[0]
    assert res.is_column[0] == 0 and res.likelihoods.sum() == 0
    assert tuple(res.genotype) == (0, 0) and res.quality[0] == 10000  # graph.cpp:225-227
    b = pg.PanelBuilder()
    v = b.add_variant(100, [0, 1]); b.set_coverage(v, 20)  # a column without unique k-mers
    p = b.build()
    want = oracles.cpu_hmm_run(oracle, "pgo_", [p], table)[0]
    got = engine.hmm_run([p], table)[0]
    assert_results_close(got, want)


def test_full_size_properties(engine):
    """cfg-scale shape (H=64) without an oracle: rows are normalised, GT is the argmax, results are
    reproducible and independent of how chromosomes are batched."""
    from synthdata import small as synth
    wl = synth.make_workload(n_chrom=3, n_variants=6000, n_haplotypes=64, coverage=0, with_reads=False, seed=4)
    synth.fill_synthetic_counts(np.random.default_rng(4), wl)
    table = pg.ProbabilityTable(6, 96, 48, 0.01)
    r1 = engine.hmm_run(wl.panels, table, recombrate=1.26, effective_N=1e-5)
    for r in r1:
        sums = np.add.reduceat(r.likelihoods, r.gl_offsets[:-1].astype(np.int64))
        assert np.allclose(sums[r.is_column == 1], 1.0, atol=1e-9)
        assert (r.is_column == 1).all()
    r2 = [engine.hmm_run([p], table, recombrate=1.26, effective_N=1e-5)[0] for p in wl.panels]
    for a, b in zip(r1, r2):
        assert np.array_equal(a.likelihoods, b.likelihoods)  # bitwise: batching must not change arithmetic
    # most genotypes of the simulated sample are recovered (sanity of the whole model, not a parity claim)
    ok = tot = 0
    for r, tr, p in zip(r1, wl.truth, wl.panels):
        gt = r.genotype.reshape(-1, 2)
        t = np.sort(tr.astype(np.int16), axis=1)
        ok += int((gt == t).all(axis=1).sum()); tot += len(t)
    assert ok / tot > 0.9


def test_reference_catch_suite_passes_over_the_gpu_library():
    """The reference's own Catch tests for the hot path with HMM = integration/hmm_binding.cpp over pg_hmm_run."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_tests_gpu")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_tests_gpu not built (needs the reference tree at build time)")
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "pangenie_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    out = subprocess.run([exe, "~[Histogram*]", "~Histogram*"], capture_output=True, text=True, env=env).stdout
    tail = [l for l in out.splitlines() if l.startswith("assertions:")]
    assert tail, out[-2000:]
    # every assertion but the Viterbi haplotype one (HMMTest.cpp:438, phasing is out of scope)
    assert " 1 failed" in tail[-1] and "HMMTest.cpp:438" in out, out[-3000:]
    assert out.count("FAILED:") == 1


@pytest.mark.parametrize("n_paths", [1, 2, 3, 8, 9, 10, 16, 17, 20, 21, 29, 30, 33, 34, 35, 36, 37, 40, 41, 63, 64, 65, 68, 69, 72, 73, 136, 137, 256])
def test_path_counts_at_kernel_configuration_boundaries(engine, oracle, n_paths):
    """Every tile configuration of the chain kernels at its smallest and largest path count (csrc/genotype.cu pick_cfg),
    several checkpoint blocks each, multi-allelic and undefined alleles included."""
    rng = np.random.default_rng(4000 + n_paths)
    n_var = 150 if n_paths <= 40 else 80 if n_paths <= 140 else 20
    panel = random_panel(rng, n_var, n_paths, max_alleles=min(5, n_paths + 1), undefined_frac=0.1, ref_only_frac=0.02)
    table = _table()
    for normalize in (True, False):
        kw = dict(recombrate=1.26, effective_N=25000.0 if n_paths % 2 else 1e-5, normalize=normalize)
        want = oracles.cpu_hmm_run(oracle, "pgo_", [panel], table, **kw)[0]
        got = engine.hmm_run([panel], table, **kw)[0]
        assert_results_close(got, want, atol=1e-300, label=f"P={n_paths} normalize={normalize}")


@pytest.mark.parametrize("n_paths", [21, 33, 36, 37, 65, 68])
def test_cluster_checkpoint_walk_is_bitwise_identical(engine, oracle, monkeypatch, n_paths):
    """The checkpoint walk split over a thread-block cluster (rows over 2 or 4 CTAs, row sums exchanged through distributed
    shared memory, one cluster barrier per column) computes every cell by the same expression from the same row sums as
    the single-CTA walk: identical bits; multi-allelic columns and columns whose total underflows included."""
    rng = np.random.default_rng(800 + n_paths)
    panel = random_panel(rng, 700, n_paths, max_alleles=3, undefined_frac=0.05, ref_only_frac=0.02, kmers_per_allele=(0, 8))
    probs = pg.ProbabilityTable(0, 1, 21, 0.0)   # exact zeros: some column totals are exactly zero (uniform replacement)
    probs.modify_probability(0, 10, 0.0, 1.0, 0.0)
    probs.modify_probability(0, 20, 0.0, 0.0, 1.0)
    probs.modify_probability(0, 0, 1.0, 0.0, 0.0)
    b = pg.PanelBuilder()
    pos = 1000
    for i in range(400):
        pos += int(rng.integers(100, 2000))
        al = rng.integers(0, 2, size=n_paths)
        al[0], al[1] = 0, 1
        v = b.add_variant(pos, al)
        c = [(10, 10), (20, 0), (0, 20), (10, 0)][int(rng.integers(0, 4))]
        b.insert_kmer(v, c[0], [0]); b.insert_kmer(v, c[1], [1])
    dead = b.build()
    monkeypatch.setenv("PG_SKELETON_TILE", "0")   # the cluster walk shares its step with the generic walk, not with the lean one
    for p, t, kw in ((panel, _table(), dict(recombrate=1.26, effective_N=1e-5)), (dead, probs, dict(recombrate=1.26, effective_N=25000.0))):
        monkeypatch.setenv("PG_SKELETON_CLUSTER", "0")
        one = engine.hmm_run([p, p], t, **kw)
        want = oracles.cpu_hmm_run(oracle, "pgo_", [p], t, **kw)[0]
        assert_results_close(one[0], want, atol=1e-300, label=f"single CTA P={n_paths}")
        for C in ("2", "4"):
            monkeypatch.setenv("PG_SKELETON_CLUSTER", C)
            got = engine.hmm_run([p, p], t, **kw)
            for g, o in zip(got, one):
                assert np.array_equal(g.likelihoods, o.likelihoods) and np.array_equal(g.genotype, o.genotype), (n_paths, C)


@pytest.mark.parametrize("n_paths", [17, 25, 33, 34, 35, 50, 65, 68])
def test_lean_checkpoint_walk_matches_generic_walk_and_oracle(engine, oracle, monkeypatch, n_paths):
    """The lean checkpoint walk (TMA descriptor ring + mbarriers, csrc/hmm_kernels.cuh skeleton_lean_kernel; the default for
    16 < P <= 68) against the generic walk (PG_SKELETON_TILE=0) and the oracle: multi-allelic columns (generic step inside the lean
    kernel), undefined alleles, columns whose total is exactly zero (uniform replacement, branch-free in the lean step),
    several chromosomes of different lengths in one call (chains shorter than the descriptor ring included)."""
    rng = np.random.default_rng(900 + n_paths)
    panels = [random_panel(rng, n, n_paths, max_alleles=3, undefined_frac=0.05, ref_only_frac=0.02, kmers_per_allele=(0, 8))
              for n in (700, 150, 9, 3)]
    probs = pg.ProbabilityTable(0, 1, 21, 0.0)   # exact zeros: some column totals are exactly zero
    probs.modify_probability(0, 10, 0.0, 1.0, 0.0)
    probs.modify_probability(0, 20, 0.0, 0.0, 1.0)
    probs.modify_probability(0, 0, 1.0, 0.0, 0.0)
    b = pg.PanelBuilder()
    pos = 1000
    for i in range(400):
        pos += int(rng.integers(100, 2000))
        al = rng.integers(0, 2, size=n_paths)
        al[0], al[1] = 0, 1
        v = b.add_variant(pos, al)
        c = [(10, 10), (20, 0), (0, 20), (10, 0)][int(rng.integers(0, 4))]
        b.insert_kmer(v, c[0], [0]); b.insert_kmer(v, c[1], [1])
    dead = [b.build()]
    for ps, t, kw, blk in ((panels, _table(), dict(recombrate=1.26, effective_N=1e-5), None),
                           (panels, _table(), dict(recombrate=1.26, effective_N=1e-5), "3"),     # many checkpoints, chains of 1-3 blocks
                           (dead, probs, dict(recombrate=1.26, effective_N=25000.0), "7")):
        if blk is None:
            monkeypatch.delenv("PG_HMM_B", raising=False)
        else:
            monkeypatch.setenv("PG_HMM_B", blk)
        want = oracles.cpu_hmm_run(oracle, "pgo_", ps, t, **kw)
        monkeypatch.setenv("PG_SKELETON_TILE", "0")
        generic = engine.hmm_run(ps, t, **kw)
        monkeypatch.setenv("PG_SKELETON_TILE", "3")
        lean = engine.hmm_run(ps, t, **kw)
        for g, l, w in zip(generic, lean, want):
            assert_results_close(l, w, atol=1e-300, label=f"lean P={n_paths}")
            assert_results_close(g, w, atol=1e-300, label=f"generic P={n_paths}")
            assert np.array_equal(g.genotype, l.genotype)
            np.testing.assert_allclose(l.likelihoods, g.likelihoods, rtol=1e-9, atol=1e-300)


@pytest.mark.parametrize("n_paths", [9, 33])
def test_samples_batched_in_one_call_equal_separate_calls(engine, oracle, n_paths):
    """SURVEY.md 8f row 4 (many samples on the same index, reference README.md:128): the panels of S samples go into ONE pg_hmm_run
    call as further chromosomes - every sample's forward / backward checkpoint walks run concurrently and the block kernel pulls
    S times the jobs.  Same structure, each sample its own counts and coverage: the batched results are bit-identical to the
    per-sample calls, and match the oracle."""
    rng = np.random.default_rng(4200 + n_paths)
    base = [random_panel(rng, n, n_paths, max_alleles=3, undefined_frac=0.03, kmers_per_allele=(0, 8)) for n in (500, 260, 40)]
    samples = []
    for s in range(3):
        ps = [copy.deepcopy(p) for p in base]
        for p in ps:
            p.kmer_counts[:] = rng.integers(0, 40, size=len(p.kmer_counts)).astype(np.uint16)
            p.coverage[:] = rng.integers(8, 30, size=len(p.coverage)).astype(np.uint16)
        samples.append(ps)
    table = _table()
    kw = dict(recombrate=1.26, effective_N=1e-5)
    batched = engine.hmm_run([p for ps in samples for p in ps], table, **kw)
    k = 0
    for ps in samples:
        alone = engine.hmm_run(ps, table, **kw)
        want = oracles.cpu_hmm_run(oracle, "pgo_", ps, table, **kw)
        for a, w in zip(alone, want):
            b = batched[k]
            k += 1
            assert np.array_equal(a.likelihoods, b.likelihoods) and np.array_equal(a.genotype, b.genotype) and np.array_equal(a.quality, b.quality)
            assert_results_close(b, w, atol=1e-300, label=f"batched P={n_paths}")


@pytest.mark.parametrize("n_paths", [9, 33])
def test_hmm_run_samples_shares_the_structure_and_uses_each_samples_table(engine, oracle, n_paths):
    """pg_hmm_run_samples: one host copy of the panel structure, per sample its counts, coverages and its OWN ProbabilityTable
    (every sample has its own k-mer coverage peak, src/commands.cpp:840-846).  Equal to one pg_hmm_run call per sample, bit for bit,
    and to the oracle."""
    rng = np.random.default_rng(4300 + n_paths)
    base = [random_panel(rng, n, n_paths, max_alleles=3, undefined_frac=0.03, kmers_per_allele=(0, 8)) for n in (400, 130)]
    peaks = [12, 30, 21]
    tables = [pg.ProbabilityTable(pk // 4, pk * 4, 2 * pk, 0.01) for pk in peaks]
    counts, covs = [], []
    for pk in peaks:
        counts.append([rng.integers(0, 2 * pk, size=len(p.kmer_counts)).astype(np.uint16) for p in base])
        covs.append([rng.integers(max(pk // 2, 1), 2 * pk, size=p.n_variants).astype(np.uint16) for p in base])
    kw = dict(recombrate=1.26, effective_N=1e-5)
    batched = engine.hmm_run_samples(base, counts, covs, tables, **kw)
    for s_, pk in enumerate(peaks):
        ps = [copy.deepcopy(p) for p in base]
        for p, c_, v_ in zip(ps, counts[s_], covs[s_]):
            p.kmer_counts[:] = c_
            p.coverage[:] = v_
        alone = engine.hmm_run(ps, tables[s_], **kw)
        want = oracles.cpu_hmm_run(oracle, "pgo_", ps, tables[s_], **kw)
        for a, b, w in zip(alone, batched[s_], want):
            assert np.array_equal(a.likelihoods, b.likelihoods) and np.array_equal(a.genotype, b.genotype) and np.array_equal(a.quality, b.quality)
            assert np.array_equal(a.coverage, b.coverage)
            assert_results_close(b, w, atol=1e-300, label=f"samples P={n_paths} peak={pk}")
