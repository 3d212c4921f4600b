"""Histogram smoothing + peak choice (reference src/histogram.cpp:41-63, src/sequenceutils.cpp:42-84) against the
reference's four real 10k-bin histograms and the peaks its HistogramTest.cpp:32-71 expects (56, 26, 60, 42).
tests/golden/histo/*.histo.gz are gzip copies of the reference's tests/data/test{,2,3,4}.histo (data files)."""
import ctypes as C
import gzip
import os

import numpy as np
import pytest

from pangenie_b200 import capi

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "histo")
CASES = [("test.histo.gz", 56), ("test2.histo.gz", 26), ("test3.histo.gz", 60), ("test4.histo.gz", 42)]


def _load(name):
    bins = np.zeros(10001, np.uint64)   # Histogram(filename, 10000): "count value" pairs, counts above max dropped
    with gzip.open(os.path.join(G, name), "rt") as f:
        for line in f:
            t = line.split()
            if len(t) >= 2 and t[0].isdigit() and int(t[0]) <= 10000:
                bins[int(t[0])] = int(t[1])
    return bins


@pytest.mark.parametrize("name,expected", CASES)
def test_peak_matches_reference_expectation(name, expected, oracle):
    bins = _load(name)
    lib = capi.load()
    for f in (lib.pg_histogram_peak, oracle.pgo_histogram_peak):
        b = bins.copy()
        pk = C.c_uint64(0)
        assert f(b.ctypes.data, len(b), 1, C.byref(pk)) == 0
        assert pk.value == expected


def test_simple_histogram_vectors():
    # reference tests/HistogramTest.cpp:8-30: values 1,1,1,1,5 -> single peak at 1 with height 4 before smoothing
    lib = capi.load()
    b = np.zeros(11, np.uint64)
    b[1], b[5] = 4, 1
    pk = C.c_uint64(0)
    assert lib.pg_histogram_peak(b.ctypes.data, len(b), 1, C.byref(pk)) == 0
    flat = np.zeros(11, np.uint64)
    assert lib.pg_histogram_peak(flat.ctypes.data, len(flat), 1, C.byref(pk)) != 0   # no peak: the reference throws
