"""HaplotypeSampler on the device (pg_haplotype_sample) against the reference's own class compiled unmodified
(pgr_haplotype_sample, oracle/_ref) and its CPU restatement: integer work, everything must be identical."""
import numpy as np
import pytest

import pangenie_b200 as pg
from tests import oracles
from tests.helpers import random_panel

pytestmark = pytest.mark.gpu
WHAT = ("paths", "scores", "path_to_allele", "kmer counts per variant", "counts")


@pytest.mark.parametrize("seed", range(8))
def test_sampler_matches_reference_class(oracle, seed):
    ref = oracles.load_ref()
    lib, prefix = (ref, "pgr_") if ref is not None else (oracle, "pgo_")
    rng = np.random.default_rng(300 + seed)
    n_paths = int([2, 3, 8, 20, 33, 64, 65, 129][seed])
    panel = random_panel(rng, int(rng.integers(2, 300)), n_paths, max_alleles=int(rng.choice([2, 2, 4])),
                         undefined_frac=0.1 if seed % 2 else 0.0, shared_kmer_frac=0.3, ref_only_frac=0.1,
                         kmers_per_allele=(0, 6), count_range=(0, 12), spacing=(50, 200000))   # (the reference's KmerPath16 holds 16 k-mers)
    for size, add_ref, penalty, eff_n in ((1, False, 10, 25000.0), (min(n_paths, 5), True, 5, 0.01), (min(n_paths, 15), False, 10, 1e-5)):
        want = oracles.cpu_haplotype_sample(lib, prefix, panel, size, effective_N=eff_n, add_reference=add_ref, allele_penalty=penalty)
        got = pg.haplotype_sample(panel, size, effective_N=eff_n, add_reference=add_ref, allele_penalty=penalty)
        for x, y, what in zip(got, want, WHAT):
            assert np.array_equal(x, y), (what, n_paths, size, add_ref)


def test_sampler_large_panel_and_saturation(oracle):
    """More paths than a warp has lanes x 8 (several registers per lane), every path sampled (the last passes see columns
    whose free paths have saturated costs), long chromosomes."""
    rng = np.random.default_rng(5)
    for n_paths, n_var, size in ((300, 400, 12), (40, 3000, 40), (1000, 60, 3)):
        panel = random_panel(rng, n_var, n_paths, max_alleles=3, undefined_frac=0.05, kmers_per_allele=(0, 6), count_range=(0, 12))
        want = oracles.cpu_haplotype_sample(oracle, "pgo_", panel, size, effective_N=1e-5, add_reference=True, allele_penalty=10)
        got = pg.haplotype_sample(panel, size, effective_N=1e-5, add_reference=True, allele_penalty=10)
        for x, y, what in zip(got, want, WHAT):
            assert np.array_equal(x, y), (what, n_paths, size)


def test_sampler_argument_checks():
    b = pg.PanelBuilder()
    b.add_variant(10, [0, 1])
    with pytest.raises(pg.PgError):
        pg.haplotype_sample(random_panel(np.random.default_rng(1), 3, 1100), 2)   # more than 1024 paths
    assert pg.haplotype_sample(b.build(), 0)[1].size == 0                          # size < 1: nothing to do
