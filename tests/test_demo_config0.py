"""BASELINE.json configs[0]: the reference's demo (`PanGenie-index -r demo/test-reference.fa -v demo/test-variants.vcf -o pre`
then `PanGenie -f pre -i demo/test-reads.fa`, README.md:270-271) end to end against the VCF the reference ships as the
expected result (demo/test_genotyping.vcf; byte copies of the four demo files under tests/golden/demo/).

GT / GQ / GL / KC / UK of the four records only come out right if every stage does: graph sequences and path segments
(test helper tests/refgraph.py), graph k-mer counting, unique-k-mer selection (index stage), read k-mer
counting (FASTA), histogram peak, count fill + local coverage, emissions, forward-backward, genotype + quality.
The CPU test runs the oracle chain (pins the restatements against the reference's real output); the GPU test runs the product.
GL is printed by the reference with 4 significant digits (src/graph.cpp:262-266); x87 long double resolves 1 - 8.7e-19 where
fp64 gives exactly 1, so a likelihood of log10 = -3.8e-19 may come out as 0 and its GQ (180 = quantisation of the 64-bit
mantissa) is only required to be at least 150."""
import math
import os

import numpy as np
import pytest

import pangenie_b200 as pg
from pangenie_b200.panel import Result
from tests import oracles, refgraph

D = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "demo")


def expected_records():
    recs = []
    for line in open(os.path.join(D, "test_genotyping.vcf")):
        if line.startswith("#"):
            continue
        t = line.rstrip("\n").split("\t")
        info = dict(x.split("=") for x in t[7].split(";"))
        gt, gq, gl, kc = t[9].split(":")
        recs.append(dict(pos=int(t[1]) - 1, uk=int(info["UK"]), gt=tuple(int(x) for x in gt.split("/")), gq=int(gq),
                         gl=[float(x) for x in gl.split(",")], kc=int(kc)))
    return recs


def demo_inputs():
    g = refgraph.graph_from_vcf(os.path.join(D, "test-variants.vcf"), os.path.join(D, "test-reference.fa"), 31, True)["chr1"]
    reads = open(os.path.join(D, "test-reads.fa"), "rb").read()
    return g, g.segments_fasta().encode(), reads


def check_against_vcf(res: Result, panel):
    want = expected_records()
    assert panel.n_variants == len(want) == 4
    for v, w in enumerate(want):
        assert int(panel.positions[v]) == w["pos"]
        assert int(res.unique_kmers[v]) == w["uk"], (v, res.unique_kmers[v], w["uk"])
        assert int(res.coverage[v]) == w["kc"], (v, res.coverage[v], w["kc"])
        assert tuple(int(x) for x in res.genotype[2 * v:2 * v + 2]) == w["gt"], v
        row = res.row(v)
        assert len(row) == len(w["gl"])
        for x, e in zip(row, w["gl"]):
            got = math.log10(x) if x > 0 else -math.inf
            assert abs(got - e) <= max(1e-6, 1.5e-3 * abs(e)), (v, got, e)     # 4 printed digits
        if w["gq"] < 150:
            assert abs(int(res.quality[v]) - w["gq"]) <= 1, (v, res.quality[v], w["gq"])
        else:
            assert int(res.quality[v]) >= 150, (v, res.quality[v], w["gq"])


def test_oracle_chain_reproduces_the_demo_vcf():
    lib = oracles.load_oracle()
    g, segments, reads = demo_inputs()
    graph_counts = oracles.OracleCounter(lib)
    graph_counts.feed(segments, pg.PG_OP_COUNT)
    panel = oracles.oracle_unique_kmers(lib, graph_counts, refgraph.flatten(g))
    assert panel.n_paths == 25
    counts = oracles.OracleCounter(lib, reads, segments, 31)
    peak = counts.computeHistogram(10000, True)
    counts.fill_counts(peak, [panel])
    table = pg.ProbabilityTable(peak // 4, peak * 4, 2 * peak, 0.01)
    res = oracles.cpu_hmm_run(lib, "pgo_", [panel], table, recombrate=1.26, effective_N=1e-5)[0]
    check_against_vcf(res, panel)


@pytest.mark.gpu
def test_device_path_reproduces_the_demo_vcf():
    g, segments, reads = demo_inputs()
    graph_counts = pg.KmerCounter(max_distinct=1 << 17)
    graph_counts.feed(segments, pg.PG_OP_COUNT)
    panel = pg.UniqueKmerSelection(graph_counts, refgraph.flatten(g)).panel()
    eng = pg.Engine(0)
    (res,), peak = eng.genotype_run(reads, segments, [panel], k=31, recombrate=1.26, effective_N=1e-5)
    assert peak > 0
    check_against_vcf(res, panel)
