"""Parity at BASELINE.json sizes and shapes (SNPs + indels + tri-allelic + undefined alleles, synthdata/large.py):
configs[1] at full size against the whole oracle pipeline, the configs[2] / configs[3] per-column shapes against the
reference's own hmm.cpp (oracle/_ref), streaming ingest across the staging-ring wrap-around, the canonical key layout."""
import os

import numpy as np
import pytest

import pangenie_b200 as pg
from synthdata import large
from tests import oracles
from tests.helpers import assert_results_close

pytestmark = pytest.mark.gpu
KW = dict(recombrate=1.26, effective_N=1e-5)


def _oracle_counts(oracle, wl, reads, segs):
    thr = os.cpu_count() or 1
    o = oracles.OracleCounter(oracle, None, None, wl.k)
    o.feed(segs, pg.PG_OP_PRIME, threads=thr)
    o.feed(reads, pg.PG_OP_UPDATE, threads=thr)
    return o


def _check_pipeline(engine, oracle, cpu_lib, prefix, spec, hmm_threads=1):
    wl = large.make_workload(spec, "cuda")
    reads, segs = wl.reads.cpu().numpy(), wl.segments.cpu().numpy()
    got, peak = engine.genotype_run(reads, segs, wl.panels, k=wl.k, **KW)
    counts = [(p.kmer_counts.copy(), p.coverage.copy()) for p in wl.panels]
    o = _oracle_counts(oracle, wl, reads, segs)
    assert o.computeHistogram(10000, True) == peak
    o.fill_counts(peak, wl.panels)
    for p, (c, cv) in zip(wl.panels, counts):
        assert np.array_equal(p.kmer_counts, c) and np.array_equal(p.coverage, cv)       # integer work: bit-exact
    table = pg.ProbabilityTable(peak // 4, peak * 4, 2 * peak, 0.01)
    want = oracles.cpu_hmm_run(cpu_lib, prefix, wl.panels, table, threads=hmm_threads, **KW)
    for i, (g, w) in enumerate(zip(got, want)):
        assert_results_close(g, w, label=f"chromosome {i}")                               # 1e-6 relative, identical GT
    n_multi = sum(int((np.diff(p.allele_offsets) > 2).sum()) for p in wl.panels)
    n_undef = sum(int(p.allele_undefined.sum()) for p in wl.panels)
    assert n_multi > 0 and n_undef > 0, "the workload must exercise multi-allelic and undefined-allele columns"
    # the resident path (device text, partitioned counting where the table is large) gives the same bits
    engine.load(wl.panels)
    assert engine.run_resident(wl.reads, wl.segments, k=wl.k, **KW) == peak
    for g, w in zip(engine.fetch(), got):
        assert np.array_equal(g.likelihoods, w.likelihoods) and np.array_equal(g.genotype, w.genotype)
    return wl, got


def test_configs1_full_size_against_the_oracle(engine, oracle):
    """BASELINE.json configs[1] (1 chromosome, 10k variants, 8 haplotypes, 10x = 126 MB of FASTQ: the 4 x 16 MiB staging ring wraps
    twice) through pg_genotype_run against the complete oracle pipeline."""
    wl, got = _check_pipeline(engine, oracle, oracle, "pgo_", large.CONFIGS["cfg2"])
    assert wl.reads.numel() > 64 << 20
    ok = sum(int((r.genotype.reshape(-1, 2) == np.sort(t.astype(np.int16), axis=1)).all(axis=1).sum()) for r, t in zip(got, wl.truth))
    assert ok / wl.n_variants > 0.9   # the simulated sample is recovered


@pytest.mark.parametrize("name,n_var", [("cfg3", 6000), ("cfg4", 3000)])
def test_baseline_shapes_against_reference_hmm(engine, oracle, ref, name, n_var):
    """The per-column shapes of configs[2] (P = 33) and configs[3] (P = 65), 3 chromosomes at 30x, whole pipeline; emission +
    forward-backward checked against the reference's own hmm.cpp (pgr_hmm_run_mt)."""
    spec = large.scaled(large.CONFIGS[name], n_var)
    spec = large.Spec(3, n_var, spec.n_haplotypes, spec.coverage, seed=spec.seed)
    _check_pipeline(engine, oracle, ref, "pgr_", spec, hmm_threads=3)


def test_count_create_streams_files_through_the_ring(oracle, tmp_path):
    """pg_count_create(path): file -> pinned ring -> device, more than 64 MiB so every staging buffer is reused."""
    wl = large.make_workload(large.Spec(2, 6000, 4, 10.0, seed=77), "cuda", with_panels=True)
    reads, segs = wl.reads.cpu().numpy(), wl.segments.cpu().numpy()
    assert len(reads) > 70 << 20
    rp, sp = tmp_path / "reads.fq", tmp_path / "segments.fa"
    reads.tofile(rp)
    segs.tofile(sp)
    g = pg.KmerCounter(str(rp), str(sp), wl.k)
    o = _oracle_counts(oracle, wl, reads, segs)
    codes = np.concatenate([p.kmer_codes for p in wl.panels] + [p.flank_codes for p in wl.panels])
    assert np.array_equal(g.lookup(codes), o.lookup(codes))
    assert np.array_equal(g.histogram(10000), o.histogram(10000)) and g.distinct() == o.distinct()
    # count-all mode from a file, and the reference's input checks (src/commands.cpp:42-56)
    small = reads[: (4 << 20) // wl.record_bytes * wl.record_bytes]
    small.tofile(tmp_path / "small.fq")
    g2 = pg.KmerCounter(str(tmp_path / "small.fq"), None, wl.k, hash_size=8_000_000)
    o2 = oracles.OracleCounter(oracle, small, None, wl.k)
    assert np.array_equal(g2.histogram(10000), o2.histogram(10000))
    with pytest.raises(pg.PgError):
        pg.KmerCounter(str(tmp_path / "missing.fq"), None, wl.k)
    with pytest.raises(pg.PgError):
        pg.KmerCounter(str(tmp_path / "reads.fq.gz"), None, wl.k)


def test_canonical_layout_is_independent_of_the_insertion_order(oracle):
    """PRIME the same k-mer set from differently ordered segment files, canonicalize: the bucket arrays are identical
    byte for byte (what lets every GPU prime for itself and the counts be all-reduced), and lookups still work."""
    import torch
    from pangenie_b200.distributed import _CudaArray
    wl = large.make_workload(large.Spec(2, 3000, 4, 3.0, seed=78), "cuda")
    segs = wl.segments.cpu().numpy()
    starts = np.flatnonzero(np.concatenate([[True], (segs[1:] == ord(">")) & (segs[:-1] == 10)]))
    ends = np.concatenate([starts[1:], [len(segs)]])
    rng = np.random.default_rng(3)
    perm = rng.permutation(len(starts))
    shuffled = np.concatenate([segs[starts[i]:ends[i]] for i in perm])
    reads = wl.reads.cpu().numpy()
    tabs, counters = [], []
    for text in (segs, shuffled, segs[::1].copy()):
        c = pg.KmerCounter(None, None, wl.k, max_distinct=len(segs))
        c.feed(text, pg.PG_OP_PRIME)
        c.canonicalize()
        sp, _cp, cap = c.device_arrays()
        tabs.append(torch.as_tensor(_CudaArray(sp, 2 * cap, "<i8"), device="cuda").clone())
        counters.append(c)
    assert torch.equal(tabs[0], tabs[1]) and torch.equal(tabs[0], tabs[2])
    # lookups and counting on the canonical table are unaffected
    c = counters[1]
    c.feed(reads, pg.PG_OP_UPDATE)
    o = _oracle_counts(oracle, wl, reads, segs)
    codes = np.concatenate([p.kmer_codes for p in wl.panels] + [p.flank_codes for p in wl.panels])
    assert np.array_equal(c.lookup(codes), o.lookup(codes))
    assert np.array_equal(c.histogram(10000), o.histogram(10000))
    # the exchange in pieces: export / import ranges reproduce the counts
    cap = c.capacity()
    buf_slots = 1 << 16
    addr = c.exchange_buffer(buf_slots)
    view = torch.as_tensor(_CudaArray(addr, buf_slots, "<i4"), device="cuda")
    total = 0
    for first in range(0, cap, buf_slots):
        n = min(buf_slots, cap - first)
        c.export_range(first, n)
        total += int(view[:n].sum().item())
        view[:n] *= 2
        torch.cuda.synchronize()   # the alias tensor lives on torch's stream, the import kernel on the counter's
        c.import_range(first, n)
    assert total == int(o.histogram(1 << 20).dot(np.arange((1 << 20) + 1, dtype=np.uint64)))
    assert np.array_equal(c.lookup(codes), 2 * o.lookup(codes))


def test_fasta_with_long_headers_and_fastq_layout_errors(oracle):
    """FASTA headers far longer than the 128-byte look-ahead (the '>' closes the window), and FASTQ files that are not in the
    4-line layout are rejected instead of counting quality characters."""
    rng = np.random.default_rng(9)
    recs = []
    for i in range(4000):
        hdr = ">" + "".join(rng.choice(list("abcdefgh ACGT:/|"), size=int(rng.integers(150, 400))))
        seq = "".join(rng.choice(list("ACGT"), size=int(rng.integers(20, 400))))
        lines = [seq[j:j + 70] for j in range(0, len(seq), 70)]
        recs.append(hdr + "\n" + "\n".join(lines) + "\n")
    fa = np.frombuffer("".join(recs).encode(), np.uint8)
    for k in (31, 15):
        g = pg.KmerCounter(fa, None, k, hash_size=4_000_000)
        o = oracles.OracleCounter(oracle, fa, None, k)
        assert g.distinct() == o.distinct() and np.array_equal(g.histogram(10000), o.histogram(10000))
    good = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, bytes(rng.choice(list(b"ACGT"), size=100).tolist()), b"F" * 100) for i in range(3000))
    pg.KmerCounter(np.frombuffer(good, np.uint8), None, 31, hash_size=1_000_000)
    wrapped = b"".join(b"@r%d\n%s\n%s\n+\n%s\n" % (i, b"ACGT" * 20, b"ACGT" * 5, b"A" * 100) for i in range(3000))   # sequence on two lines
    cut = good.index(b"\n@", 5000) + 1
    blank = good[:cut] + b"\n" + good[cut:]                                                                              # a stray blank line
    for bad in (wrapped, blank):
        with pytest.raises(pg.PgError) as e:
            pg.KmerCounter(np.frombuffer(bad, np.uint8), None, 31, hash_size=1_000_000)
        assert "4-line layout" in str(e.value)   # PG_ERR_FORMAT (the constructor reports the message of the failed feed)


def test_histogram_from_concurrent_host_threads(oracle):
    import threading
    wl = large.make_workload(large.Spec(1, 2000, 4, 6.0, seed=79), "cuda")
    reads, segs = wl.reads.cpu().numpy(), wl.segments.cpu().numpy()
    g = pg.KmerCounter(reads, segs, wl.k)
    want = _oracle_counts(oracle, wl, reads, segs).histogram(10000)
    bad = []

    def work():
        for _ in range(20):
            if not np.array_equal(g.histogram(10000), want):
                bad.append(1)
            g.computeKmerCoverage(1000)
    ths = [threading.Thread(target=work) for _ in range(6)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not bad
