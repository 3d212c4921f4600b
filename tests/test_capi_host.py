"""C-ABI surface and pure-host logic (no GPU needed)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import pangenie_b200 as pg
from pangenie_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = capi.load()
    hdr = open(os.path.join(ROOT, "include", "pangenie_b200.h")).read()
    declared = set(re.findall(r"\b(pg_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/pangenie_b200.h but not exported"
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    assert b"sm_100a" in lib.pg_version()


def test_no_gpu_means_loud_failure_not_fallback():
    lib = capi.load()
    if lib.pg_device_count() > 0:
        pytest.skip("a GPU is present")
    assert not lib.pg_engine_create(0)
    assert b"no CPU fallback" in lib.pg_last_error()
    with pytest.raises(pg.PgError):
        pg.KmerCounter(b">a\nACGT\n", None, 3)


def test_probability_table_matches_reference_vectors():
    # reference tests/ProbabilityTableTest.cpp:12-26 and tests/CopyNumberTest.cpp:44-58 (hand-computed values)
    t = pg.ProbabilityTable(5, 7, 11, 0.0)
    import math
    p = t.get_probability(5, 3)
    assert p[0] == pytest.approx(0.99 * 0.01 ** 3, rel=1e-12)
    assert p[1] == pytest.approx(math.exp(-2.5) * 2.5 ** 3 / 6, rel=1e-12)
    assert p[2] == pytest.approx(math.exp(-5.0) * 5.0 ** 3 / 6, rel=1e-12)
    # outside the table: computed on the fly with the same formulas
    q = t.get_probability(12, 20)
    assert q[0] == pytest.approx(0.95 * 0.05 ** 20, rel=1e-10)
    # regularised CopyNumber: p0=(c0+r)/s, p1=(c1+r)/s, p2 = 1-p0-p1
    tr = pg.ProbabilityTable(5, 7, 11, 0.01)
    c0, c1, c2 = p
    s = c0 + c1 + c2 + 0.03
    pr = tr.get_probability(5, 3)
    assert pr[0] == pytest.approx((c0 + 0.01) / s, rel=1e-12)
    assert pr[1] == pytest.approx((c1 + 0.01) / s, rel=1e-12)
    assert pr[2] == pytest.approx((c2 + 0.01) / s, rel=1e-12)
    t.modify_probability(5, 10, 0.1, 0.9, 0.1)
    assert t.get_probability(5, 10) == pytest.approx((0.1, 0.9, 0.1), rel=1e-15)
    with pytest.raises(pg.PgError):
        t.modify_probability(9, 10, 0.1, 0.9, 0.1)


def test_probability_table_agrees_with_oracle_and_reference(oracle, ref):
    t = pg.ProbabilityTable(4, 72, 36, 0.01)
    for cov in (4, 18, 40, 71, 72, 200):
        for count in (0, 3, 35, 36, 500):
            for cn in range(3):
                a = np.log(t.get_probability(cov, count)[cn])
                assert a == pytest.approx(oracle.pgo_log_probability(cov, count, 0.01, cn), rel=1e-13, abs=1e-13)
                assert a == pytest.approx(ref.pgr_log_probability(4, 72, 36, 0.01, cov, count, cn), rel=1e-13, abs=1e-13)


def test_result_layout_matches_python_model():
    rng = np.random.default_rng(1)
    from tests.helpers import random_panel
    p = random_panel(rng, 20, 6, max_alleles=5)
    lib = capi.load()
    off = np.zeros(p.n_variants + 1, np.uint64)
    ps = p.as_struct()
    assert lib.pg_result_layout(C.byref(ps), off.ctypes.data) == 0
    assert np.array_equal(off, p.result_layout())


def test_index_selection_reports_bad_arguments_without_touching_a_device():
    """pg_unique_kmers_compute validates its arguments before any CUDA call; errors come back as NULL +
    pg_last_error(), like the reference's runtime_error texts."""
    import ctypes as C
    from pangenie_b200.capi import PgVariants
    lib = capi.load()
    vs = PgVariants()
    assert not lib.pg_unique_kmers_compute(0, None, C.byref(vs))
    assert b"null" in lib.pg_last_error()
    assert lib.pg_unique_kmers_panel(None, None) == capi.PG_ERR_ARG
    lib.pg_unique_kmers_free(None)   # like free(NULL)
