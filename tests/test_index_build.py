"""Index stage: unique-k-mer selection of `StepwiseUniqueKmerComputer`
(reference src/stepwiseuniquekmercomputer.cpp:11-93, 95-197, 227-264).

Golden pin: the reference's own index fixture — `tests/data/index_chr1_Graph.cereal` is the graph the real PanGenie-index
serialised (2 bubbles, 44 / 45 alleles, undefined alleles, 215 paths), `index_path_segments.fasta` the file its graph
k-mers were counted from, `index_chr1_kmers.tsv.gz` + `index_UniqueKmersMap.cereal` what its StepwiseUniqueKmerComputer
produced from them (byte copies under tests/golden/counting/).  The CPU restatement (oracle/pg_oracle_index.cpp) and the device
kernel (csrc/index_build.cu) must both reproduce those outputs exactly; randomised bubbles then compare the kernel with the
restatement."""
import gzip
import os

import numpy as np
import pytest

import pangenie_b200 as pg
from pangenie_b200 import refindex
from tests import oracles, refgraph

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "counting")
FIELDS = ("positions", "path_to_allele", "coverage", "kmer_offsets", "kmer_counts", "allele_offsets", "allele_ids",
          "allele_undefined", "allele_kmer_offset", "allele_kmer_mask", "kmer_codes", "flank_offsets", "flank_codes")


def assert_same_panel(got, want):
    assert got.n_paths == want.n_paths and got.n_variants == want.n_variants
    for f in FIELDS:
        a, b = getattr(got, f), getattr(want, f)
        assert np.array_equal(a, b), (f, a[:16], b[:16])


def reference_fixture():
    g = refgraph.read_graph_cereal(os.path.join(G, "index_chr1_Graph.cereal"))
    _, panels, _ = refindex.read_unique_kmers_map(os.path.join(G, "index_UniqueKmersMap.cereal"))
    want = refindex.attach_kmers_tsv(panels["chr1"], os.path.join(G, "index_chr1_kmers.tsv.gz"))
    return g, want


def random_graph(seed, n_bubbles=40, n_paths=9, k=31, with_n=True, repeats=True):
    """Bubbles with SNPs, indels, multi-allelic sites, alleles no path carries, undefined alleles, repeated sequence (k-mers
    that are not unique in the graph or occur twice inside an allele), bubbles closer than 2k (short overhangs) and
    undefined bases in the reference."""
    rng = np.random.default_rng(seed)
    L = n_bubbles * 90 + 200
    ref = rng.choice(list("ACGT"), L)
    if repeats:                               # copy a few stretches elsewhere -> graph count 2
        for _ in range(n_bubbles // 4):
            a, b, n = rng.integers(0, L - 80), rng.integers(0, L - 80), rng.integers(20, 70)
            ref[b:b + n] = ref[a:a + n]
        for _ in range(n_bubbles // 8):       # low-complexity stretch: the same k-mer twice inside one allele
            a = rng.integers(0, L - 100)
            ref[a:a + 80] = np.resize(ref[a:a + rng.integers(1, 4)], 80)
    if with_n:
        for _ in range(3):
            ref[rng.integers(0, L)] = "N"
    ref = "".join(ref)
    g = refgraph.RefGraph("chrT", k, True, ref)
    pos = 2 * k
    for _ in range(n_bubbles):
        pos += int(rng.choice([k, k + 1, 40, 70, 90, 130]))
        if pos + 60 + k >= L:
            break
        kind = rng.random()
        reflen = 1 if kind < 0.6 else int(rng.integers(1, 20))
        refa = ref[pos:pos + reflen]
        if "N" in refa:
            continue
        n_alt = 1 if kind < 0.7 else int(rng.integers(2, 6))
        alts = []
        while len(alts) < n_alt:
            a = "".join(rng.choice(list("ACGT"), int(rng.integers(1, 30)) if kind > 0.6 else 1))
            if a != refa and a not in alts:
                alts.append(a)
        alleles = [refa] + alts
        undefined = [False] * len(alleles)
        if rng.random() < 0.15:
            alleles.append("N" * int(rng.integers(1, 4)))
            undefined.append(True)
        n_all = len(alleles)
        carried = list(range(n_all)) if rng.random() < 0.7 else [a for a in range(n_all) if a == 0 or undefined[a] or rng.random() < 0.6]
        paths = [0] + [int(rng.choice(carried)) for _ in range(n_paths - 1)]
        for a in range(n_all):               # the reference throws on an undefined allele no path carries
            if undefined[a] and a not in paths:
                paths[-1] = a
        end = pos + reflen
        left, right = ref[pos - (k - 1):pos], ref[end:end + k - 1]
        g.bubbles.append(refgraph.Bubble("chrT", pos, end, [left + a + right for a in alleles], undefined, paths))
        pos = end
    g.set_overhangs()
    return g


def test_oracle_reproduces_the_reference_index_fixture():
    lib = oracles.load_oracle()
    g, want = reference_fixture()
    segments = open(os.path.join(G, "index_path_segments.fasta"), "rb").read()
    assert g.segments_fasta().encode() == segments             # GraphBuilder::write_path_segments restated by the test helper
    counts = oracles.OracleCounter(lib)
    counts.feed(segments, pg.PG_OP_COUNT)
    got = oracles.oracle_unique_kmers(lib, counts, refgraph.flatten(g))
    assert got.n_variants == 2 and got.n_paths == 215
    assert_same_panel(got, want)


def test_oracle_on_random_bubbles_is_self_consistent():
    """Selection invariants the reference's UniqueKmerComputerTest checks (tests/UniqueKmerComputerTest.cpp:36-46, 140-150)."""
    lib = oracles.load_oracle()
    g = random_graph(5, n_bubbles=60, n_paths=33)
    counts = oracles.OracleCounter(lib)
    counts.feed(g.segments_fasta().encode(), pg.PG_OP_COUNT)
    pan = oracles.oracle_unique_kmers(lib, counts, refgraph.flatten(g))
    assert pan.n_variants == len(g.bubbles)
    for v in range(pan.n_variants):
        n = int(pan.kmer_offsets[v + 1] - pan.kmer_offsets[v])
        on = 0
        for a in range(int(pan.allele_offsets[v]), int(pan.allele_offsets[v + 1])):
            on += bin(int(pan.allele_kmer_mask[a])).count("1")
        assert on == n and n <= 301                            # every k-mer on exactly one allele
        assert int(pan.flank_offsets[v + 1] - pan.flank_offsets[v]) <= 24


@pytest.mark.gpu
def test_device_selection_reproduces_the_reference_index_fixture(tmp_path):
    g, want = reference_fixture()
    counts = pg.KmerCounter(max_distinct=1 << 16)
    counts.feed(open(os.path.join(G, "index_path_segments.fasta"), "rb").read(), pg.PG_OP_COUNT)
    sel = pg.UniqueKmerSelection(counts, refgraph.flatten(g))
    assert_same_panel(sel.panel(), want)
    # the k-mer table file, byte for byte what the real PanGenie-index wrote (after decompression)
    out = tmp_path / "index_chr1_kmers.tsv.gz"
    sel.write_tsv("chr1", [b.end for b in g.bubbles], str(out))
    assert gzip.open(out, "rb").read() == gzip.open(os.path.join(G, "index_chr1_kmers.tsv.gz"), "rb").read()
    ms, n = sel.stats()
    assert n > 0 and ms >= 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n_paths,k", [(1, 9, 31), (2, 33, 31), (3, 120, 31), (4, 9, 21), (6, 330, 32), (7, 5, 11)])
def test_device_selection_matches_oracle_on_random_bubbles(seed, n_paths, k):
    lib = oracles.load_oracle()
    g = random_graph(seed, n_bubbles=150, n_paths=n_paths, k=k)
    seg = g.segments_fasta().encode()
    flat = refgraph.flatten(g)
    oc = oracles.OracleCounter(lib, k=k)
    oc.feed(seg, pg.PG_OP_COUNT)
    want = oracles.oracle_unique_kmers(lib, oc, flat)
    counts = pg.KmerCounter(kmer_size=k, max_distinct=1 << 18)
    counts.feed(seg, pg.PG_OP_COUNT)
    got = pg.UniqueKmerSelection(counts, flat).panel()
    assert_same_panel(got, want)
    assert int(want.kmer_offsets[-1]) > 0 and int(want.flank_offsets[-1]) > 0


@pytest.mark.gpu
def test_device_selection_large_bubble_uses_the_global_scratch():
    """A bubble with more k-mers than fit the shared-memory sort (many long alleles)."""
    lib = oracles.load_oracle()
    rng = np.random.default_rng(11)
    k = 31
    ref = "".join(rng.choice(list("ACGT"), 4000))
    g = refgraph.RefGraph("chrL", k, True, ref)
    pos, reflen = 500, 40
    alleles = [ref[pos:pos + reflen]] + ["".join(rng.choice(list("ACGT"), int(rng.integers(150, 400)))) for _ in range(40)]
    left, right = ref[pos - (k - 1):pos], ref[pos + reflen:pos + reflen + k - 1]
    paths = [0] + [int(rng.integers(0, len(alleles))) for _ in range(64)]
    g.bubbles.append(refgraph.Bubble("chrL", pos, pos + reflen, [left + a + right for a in alleles], [False] * len(alleles), paths))
    g.bubbles.append(refgraph.Bubble("chrL", 2000, 2001, [ref[1970:2000] + a + ref[2001:2031] for a in (ref[2000], "A" if ref[2000] != "A" else "C")],
                                     [False, False], [0] + [int(rng.integers(0, 2)) for _ in range(64)]))
    g.set_overhangs()
    flat = refgraph.flatten(g)
    assert int(flat["seq_offsets"][len(alleles)]) - len(alleles) * (k - 1) > 4096
    seg = g.segments_fasta().encode()
    oc = oracles.OracleCounter(lib, k=k)
    oc.feed(seg, pg.PG_OP_COUNT)
    want = oracles.oracle_unique_kmers(lib, oc, flat)
    counts = pg.KmerCounter(kmer_size=k, max_distinct=1 << 18)
    counts.feed(seg, pg.PG_OP_COUNT)
    got = pg.UniqueKmerSelection(counts, flat).panel()
    assert_same_panel(got, want)
    assert int(want.kmer_offsets[1]) == 301                    # max(301, P) k-mers, round-robin over the alleles


def test_closed_form_of_the_round_robin_equals_the_reference_loop():
    """`select_kmers` lets the alleles take turns (src/stepwiseuniquekmercomputer.cpp:74-92).  The kernel's single-thread walk uses the
    closed form "rounds taken completely + the first alleles of the cut round" (csrc/index_build.cu); this mirrors both in Python
    over random queue lengths, caps and totals (round boundaries that hit the total exactly included)."""
    import random

    def reference(q, max_kmers, max_total):
        taken, left, n, keep = {a: 0 for a in q}, dict(q), 0, True
        while n < max_total and keep:
            added = False
            for a in sorted(q):
                if left[a] > 0 and taken[a] < max_kmers:
                    taken[a] += 1; left[a] -= 1; added = True; n += 1
                if n >= max_total:
                    break
            keep = added
        return taken

    def closed_form(q, max_kmers, max_total):
        rounds = [0] * 33
        for a in q:
            for j in range(min(q[a], max_kmers)):
                rounds[j] += 1
        full = taken_n = 0
        while full < max_kmers and rounds[full] > 0 and taken_n + rounds[full] <= max_total:
            taken_n += rounds[full]; full += 1
        extra = max_total - taken_n if (full < max_kmers and taken_n < max_total) else 0
        if full < max_kmers:
            extra = min(extra, rounds[full])
        taken = {a: 0 for a in q}
        for a in sorted(q):
            for j in range(q[a]):
                take = j < full
                if not take and j == full and extra > 0:
                    take, extra = True, extra - 1
                taken[a] += int(take)
        return taken

    rnd = random.Random(7)
    for _ in range(20000):
        q = {a: rnd.choice([1, 2, 3, 5, 16, 17, 31, 32, 33, 40]) for a in rnd.sample(range(50), rnd.randint(1, 12))}
        mk, mt = rnd.choice([16, 32, 3, 1]), rnd.choice([301, 5, 17, 40, 64, 33])
        assert reference(q, mk, mt) == closed_form(q, mk, mt), (q, mk, mt)
