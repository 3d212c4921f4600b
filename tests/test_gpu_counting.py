"""GPU k-mer counter vs the CPU restatement of the jellyfish semantics (bit-exact: integer work)."""
import numpy as np
import pytest

import pangenie_b200 as pg
from synthdata import small as synth
from tests import oracles

pytestmark = pytest.mark.gpu


def _codes(rng, n, k):
    return rng.integers(0, 4 ** k if k < 32 else 2 ** 63, size=n, dtype=np.uint64)


def _compare(oracle, reads, segments, k, probe_codes, hash_size=600_000):
    g = pg.KmerCounter(reads, segments, k, hash_size=hash_size)
    o = oracles.OracleCounter(oracle, reads, segments, k)
    assert np.array_equal(g.histogram(300), o.histogram(300))
    assert np.array_equal(g.lookup(probe_codes), o.lookup(probe_codes))
    assert g.distinct() == o.distinct()
    return g, o


def test_reference_kmercounter_vectors(oracle):
    # reference tests/KmerCounterTest.cpp:10-32 with tests/data/reads.fa and kmerfile.fa
    reads = b">read1\nATGCTGTAAAAAAACGGC\n"
    g = pg.KmerCounter(reads, None, 10, hash_size=1000)
    seq = "ATGCTGTAAAAAAACGGC"
    for i in range(len(seq) - 9):
        assert g.getKmerAbundance(seq[i:i + 10]) == 1
    kmerfile = b">kmers\nATGCTGTAAAA\n"
    g2 = pg.KmerCounter(reads, kmerfile, 10, hash_size=1000)
    assert g2.getKmerAbundance("ATGCTGTAAA") == 1 and g2.getKmerAbundance("TGCTGTAAAA") == 1
    for i in range(2, len(seq) - 9):
        assert g2.getKmerAbundance(seq[i:i + 10]) == 0
    # canonical: the reverse complement of a counted k-mer has the same abundance
    assert g.getKmerAbundance("TTTACAGCAT") == 1
    assert g.getKmerAbundance("ATGCTGTAAN") == 0


@pytest.mark.parametrize("k", [31, 21, 32, 5])
def test_fastq_and_fasta_match_oracle(oracle, k):
    rng = np.random.default_rng(k)
    wl = synth.make_workload(n_chrom=2, n_variants=300, n_haplotypes=4, coverage=4.0, k=31, seed=k)
    probes = np.concatenate([p.kmer_codes for p in wl.panels])[:5000] & np.uint64((1 << (2 * k)) - 1 if k < 32 else 0xFFFFFFFFFFFFFFFF)
    probes = np.concatenate([probes, _codes(rng, 2000, k)])
    _compare(oracle, wl.reads_fastq, wl.segments_fasta, k, probes)      # PRIME + UPDATE (default mode)
    _compare(oracle, wl.reads_fastq[:400_000 // wl.record_bytes * wl.record_bytes], None, k, probes)  # count-all mode


def test_edge_cases_match_oracle(oracle):
    rng = np.random.default_rng(0)
    probes = _codes(rng, 100, 7)
    cases = [
        b">a\nACGTACGTAC\n",                                   # shorter than k
        b">a\nACGTNACGTACGTACGTNNACGTACGTACGT\n>b\nacgtacgtacgtacgt\n",   # N resets, lower case
        b">a\nACGTACG\nTACGTACGT\n\nACGT\n>b\n>c\nAC\n",        # multi-line, empty line, empty records
        b"@r1\nACGTACGTACGTAAAA\n+\n@@@@++++IIIIFFFF\n@r2\nTTTTTTTTTTTTTTTT\n+r2\n+@+@+@+@+@+@+@+@\n",  # '@'/'+' in quality
        b"@r1\nAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA\n+\n" + b"F" * 68 + b"\n",  # homopolymer
        b">a\nACGTACGTACGT",                                   # no trailing newline
    ]
    for text in cases:
        for segs in (None, b">s\nACGTACGTACGTAAAATTTTTTTTTT\n"):
            g = pg.KmerCounter(text, segs, 7, hash_size=10_000)
            o = oracles.OracleCounter(oracle, text, segs, 7)
            all7 = np.arange(4 ** 7, dtype=np.uint64)
            assert np.array_equal(g.lookup(all7), o.lookup(all7)), (text[:30], segs)
            assert np.array_equal(g.histogram(100), o.histogram(100))


def test_tile_and_chunk_boundaries(oracle):
    """Long multi-line FASTA records and reads crossing the 8064-byte tile advance."""
    rng = np.random.default_rng(2)
    seq = synth._ASCII[rng.integers(0, 4, size=300_000)]
    fa = synth._fasta_record("chr", np.frombuffer(seq, np.uint8) if False else rng.integers(0, 4, size=300_000).astype(np.uint8), width=61)
    probes = _codes(rng, 1000, 13)
    g, o = _compare(oracle, fa, None, 13, probes)
    all_codes = np.arange(4 ** 9, dtype=np.uint64)
    g9 = pg.KmerCounter(fa, None, 9, hash_size=400_000)
    o9 = oracles.OracleCounter(oracle, fa, None, 9)
    assert np.array_equal(g9.lookup(all_codes), o9.lookup(all_codes))


def test_incremental_feed_device_and_host_agree(oracle):
    import torch
    wl = synth.make_workload(n_chrom=1, n_variants=200, n_haplotypes=4, coverage=6.0, seed=3)
    o = oracles.OracleCounter(oracle, wl.reads_fastq, wl.segments_fasta, 31)
    g = pg.KmerCounter(None, None, 31, max_distinct=len(wl.segments_fasta))
    g.feed(wl.segments_fasta, pg.PG_OP_PRIME)
    rec = wl.record_bytes  # fixed-width FASTQ records: split the reads in two record-aligned shards
    half = (len(wl.reads_fastq) // rec // 2) * rec
    g.feed(torch.from_numpy(wl.reads_fastq[:half]).cuda(), pg.PG_OP_UPDATE)   # device-resident text
    g.feed(wl.reads_fastq[half:], pg.PG_OP_UPDATE)                             # host text
    probes = np.concatenate([p.kmer_codes for p in wl.panels])
    assert np.array_equal(g.lookup(probes), o.lookup(probes))
    assert np.array_equal(g.histogram(), o.histogram())
    assert g.computeHistogram(10000, True) == o.computeHistogram(10000, True)
    assert g.computeKmerCoverage(12345) == o.computeKmerCoverage(12345)


@pytest.mark.parametrize("part_kb", [0, 16, 1])
def test_partitioned_counting_matches_oracle(oracle, monkeypatch, part_kb):
    """Large tables are counted partition by partition (csrc/kmer_count.cu: scatter + probe_parts_kernel); the knobs shrink
    the partition size so small inputs take that path too.  part_kb=0 is the direct path."""
    monkeypatch.setenv("PG_COUNT_PART_KB", str(part_kb))
    monkeypatch.setenv("PG_COUNT_PART_MIN_TEXT", "0")
    monkeypatch.setenv("PG_COUNT_PART_STAGED", "1")  # by default only device-resident text is partitioned
    wl = synth.make_workload(n_chrom=2, n_variants=300, n_haplotypes=4, coverage=4.0, k=31, seed=11)
    rng = np.random.default_rng(5)
    probes = np.concatenate([np.concatenate([p.kmer_codes for p in wl.panels])[:5000], _codes(rng, 2000, 31)])
    _compare(oracle, wl.reads_fastq, wl.segments_fasta, 31, probes)      # PRIME + UPDATE
    _compare(oracle, wl.reads_fastq[:400_000 // wl.record_bytes * wl.record_bytes], None, 31, probes)  # count-all (inserting) mode
    # several super-chunks: the partition buffers are filled and worked off more than once per pass
    monkeypatch.setenv("PG_COUNT_SUPER_MB", "1")
    _compare(oracle, wl.reads_fastq, wl.segments_fasta, 31, probes)
    monkeypatch.delenv("PG_COUNT_SUPER_MB")
    # skew: one k-mer dominates, its partition region overflows and the excess is probed directly
    poly = b"".join(b"@r%d\n" % i + b"A" * 150 + b"\n+\n" + b"F" * 150 + b"\n" for i in range(3000))
    text = poly + bytes(wl.reads_fastq[:200 * wl.record_bytes])
    g = pg.KmerCounter(text, None, 31, hash_size=600_000)
    o = oracles.OracleCounter(oracle, text, None, 31)
    assert g.getKmerAbundance("A" * 31) == 3000 * 120
    assert np.array_equal(g.lookup(probes), o.lookup(probes))
    assert g.distinct() == o.distinct()


def test_unsupported_inputs_fail_loudly():
    with pytest.raises(pg.PgError):
        pg.KmerCounter(b"ACGT\n", None, 3, hash_size=100)          # neither FASTA nor FASTQ
    with pytest.raises(pg.PgError):
        rng = np.random.default_rng(1)
        pg.KmerCounter(synth._fasta_record("a", rng.integers(0, 4, size=40_000).astype(np.uint8)), None, 11, hash_size=10)  # count-all table too small


def _random_text(rng, fastq: bool):
    """Random FASTA/FASTQ with ragged line lengths, empty records, lower case, N / IUPAC / CR bytes."""
    alphabet = np.frombuffer(b"ACGTACGTACGTACGTacgtNRYKM\r", np.uint8)
    out = bytearray()
    for r in range(int(rng.integers(1, 40))):
        n = int(rng.choice([0, 1, 5, 6, 7, 8, 15, 16, 17, 31, 64, 100, 250, 4000, 9000]))
        seq = alphabet[rng.integers(0, len(alphabet), size=n)].tobytes()
        if fastq:
            qual = bytes(rng.integers(33, 75, size=n, dtype=np.uint8))  # includes '@', '+', '>'
            out += b"@r%d some text\n" % r + seq + b"\n+" + (b"r%d" % r if rng.random() < 0.3 else b"") + b"\n" + qual + b"\n"
        else:
            out += b">rec%d\n" % r
            width = int(rng.choice([1, 7, 16, 60, 61, 4096, 100000]))
            for i in range(0, n, width):
                out += seq[i:i + width] + b"\n"
                if rng.random() < 0.05:
                    out += b"\n"  # empty line inside a record
    if rng.random() < 0.3 and not fastq and out.endswith(b"\n"):
        out = out[:-1]  # no trailing newline
    return bytes(out)


@pytest.mark.parametrize("seed", range(6))
def test_fuzzed_formats_match_oracle(oracle, seed):
    """Ragged random FASTA / FASTQ (empty records, lines of 1..100000 bases, IUPAC codes, CR bytes, '@' / '>' in quality
    strings): every 7-mer count and the histogram agree with the CPU restatement, in all three counting modes."""
    rng = np.random.default_rng(900 + seed)
    all7 = np.arange(4 ** 7, dtype=np.uint64)
    segs = _random_text(np.random.default_rng(77), False)
    for trial in range(6):
        text = _random_text(rng, fastq=bool(trial & 1))
        for s in (None, segs):
            g = pg.KmerCounter(text, s, 7, hash_size=40_000)
            o = oracles.OracleCounter(oracle, text, s, 7)
            assert np.array_equal(g.lookup(all7), o.lookup(all7)), (seed, trial, s is None)
            assert np.array_equal(g.histogram(200), o.histogram(200))
            assert g.distinct() == o.distinct()
