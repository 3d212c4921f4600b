"""Random panels and comparison helpers shared by the CPU and the GPU parity tests."""
from __future__ import annotations

import numpy as np

import pangenie_b200 as pg


def random_panel(rng, n_variants, n_paths, max_alleles=2, kmers_per_allele=(0, 6), undefined_frac=0.0,
                 shared_kmer_frac=0.0, ref_only_frac=0.1, count_range=(0, 40), cov_range=(8, 30), spacing=(50, 5000)):
    b = pg.PanelBuilder()
    pos = 1000
    for _ in range(n_variants):
        pos += int(rng.integers(*spacing))
        nall = int(rng.integers(2, max_alleles + 1))
        if rng.random() < ref_only_frac:
            alleles = np.zeros(n_paths, int)
        else:
            alleles = rng.integers(0, nall, size=n_paths)
        v = b.add_variant(pos, alleles)
        present = sorted(set(int(a) for a in alleles))
        undefined = [a for a in present if a != 0 and rng.random() < undefined_frac]
        for a in present:
            if a in undefined:
                continue  # undefined alleles carry no k-mers in the reference's index
            for _k in range(int(rng.integers(kmers_per_allele[0], kmers_per_allele[1] + 1))):
                b.insert_kmer(v, int(rng.integers(*count_range)), [a])
        for _k in range(int(rng.integers(0, 3)) if shared_kmer_frac > 0 and rng.random() < shared_kmer_frac else 0):
            on = [a for a in present if a not in undefined]
            b.insert_kmer(v, int(rng.integers(*count_range)), on[:2])
        for a in undefined:
            b.set_undefined_allele(v, a)
        b.set_coverage(v, int(rng.integers(*cov_range)))
    return b.build()


def assert_results_close(got, want, rtol=1e-6, atol=1e-12, check_gq=True, label=""):
    """Posteriors within rtol (BASELINE.json north_star: 1e-6 relative), identical GT calls."""
    assert np.array_equal(got.is_column, want.is_column), f"{label}: column sets differ"
    np.testing.assert_allclose(got.likelihoods, want.likelihoods, rtol=rtol, atol=atol, err_msg=f"{label}: likelihoods")
    assert np.array_equal(got.genotype, want.genotype), f"{label}: GT calls differ at {np.nonzero(got.genotype != want.genotype)[0][:10]}"
    assert np.array_equal(got.unique_kmers, want.unique_kmers)
    assert np.array_equal(got.coverage, want.coverage)
    if check_gq:
        # GQ = floor(-10 log10(1 - p)): allow +-1 where 1-p sits on an integer boundary within rounding
        # Above ~GQ 150 the reference's 1.0L - p is quantisation noise of the x87 grid (2^-64 = GQ 192.7), and
        # 10000 is its value for an exact 0: there only the magnitude is comparable.
        g, w = got.quality.astype(np.int64), want.quality.astype(np.int64)
        lo = w < 150
        d = np.abs(g - w)
        assert d[lo].max(initial=0) <= 1, f"{label}: GQ differs by {d[lo].max()}"
        assert (g[~lo] >= 140).all(), f"{label}: high-confidence GQ collapsed"
