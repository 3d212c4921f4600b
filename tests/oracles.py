"""Loaders for the CPU oracles (test infrastructure; the product package never imports this)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from pangenie_b200 import capi
from pangenie_b200.capi import PgHmmResult, PgPanel, ptr
from pangenie_b200.model import hmm_params
from pangenie_b200.panel import Result

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "libpg_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libpg_ref.so")

_oracle = None
_ref = None


def load_oracle():
    global _oracle
    if _oracle is None:
        src = os.path.join(ROOT, "oracle", "pg_oracle.cpp")
        if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], stdout=subprocess.DEVNULL)
        _oracle = capi.bind(C.CDLL(ORACLE_SO), "pgo_")
        _oracle.pgo_log_probability.restype = C.c_double
        _oracle.pgo_log_probability.argtypes = [C.c_uint16, C.c_uint16, C.c_double, C.c_int]
    return _oracle


def load_ref():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            if os.path.isdir("/root/reference/src"):
                subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
            else:
                return None
        _ref = capi.bind(C.CDLL(REF_SO), "pgr_")
        _ref.pgr_log_probability.restype = C.c_double
        _ref.pgr_log_probability.argtypes = [C.c_uint16, C.c_uint16, C.c_uint16, C.c_double, C.c_uint16, C.c_uint16, C.c_int]
        _ref.pgr_transitions.restype = None
        _ref.pgr_transitions.argtypes = [C.c_uint64, C.c_uint64, C.c_double, C.c_uint16, C.c_int, C.c_double, C.c_void_p]
        _ref.pgr_histogram_peak.restype = C.c_int
        _ref.pgr_histogram_peak.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]
    return _ref


def cpu_hmm_run(lib, prefix, panels, table, threads=1, **kw):
    """Runs pgo_/pgr_ hmm_run over a list of panels -> list[Result]."""
    results = [Result(p) for p in panels]
    prm, _keep = hmm_params(**kw)
    pa = (PgPanel * len(panels))()
    ra = (PgHmmResult * len(panels))()
    for i, (p, r) in enumerate(zip(panels, results)):
        pa[i] = p.as_struct()
        ra[i] = r.as_struct()
    if threads > 1:
        st = getattr(lib, prefix + "hmm_run_mt")(len(panels), pa, C.byref(table.t), C.byref(prm), ra, threads)
    else:
        st = getattr(lib, prefix + "hmm_run")(len(panels), pa, C.byref(table.t), C.byref(prm), ra)
    if st != 0:
        raise RuntimeError(getattr(lib, prefix + "last_error")().decode())
    return results


def cpu_emission_run(lib, prefix, panel, table):
    V = panel.n_variants
    off = np.zeros(V + 1, np.uint64)
    for v in range(V):
        n = panel.nr_alleles(v)
        off[v + 1] = off[v] + n * n
    em = np.zeros(int(off[-1]), np.float64)
    ls = np.zeros(V, np.float64)
    ps = panel.as_struct()
    st = getattr(lib, prefix + "emission_run")(C.byref(ps), C.byref(table.t), ptr(off), ptr(em), ptr(ls))
    if st != 0:
        raise RuntimeError(getattr(lib, prefix + "last_error")().decode())
    return off, em, ls


class OracleCounter:
    """pgo_counter wrapper with the KmerCounter method names."""

    def __init__(self, lib, reads=None, segments=None, k=31):
        self.lib, self.k = lib, k
        if reads is None:
            self.h = lib.pgo_count_new(k)
        else:
            r = np.frombuffer(reads, np.uint8) if not isinstance(reads, np.ndarray) else reads
            s = None if segments is None else (np.frombuffer(segments, np.uint8) if not isinstance(segments, np.ndarray) else segments)
            self._keep = (r, s)
            self.h = lib.pgo_count_create_from_buffers(r.ctypes.data, r.size, None if s is None else s.ctypes.data,
                                                       0 if s is None else s.size, k)
        if not self.h:
            raise RuntimeError(lib.pgo_last_error().decode())

    def feed(self, text, op, threads=1):
        a = np.frombuffer(text, np.uint8) if not isinstance(text, np.ndarray) else text
        st = self.lib.pgo_count_feed_mt(self.h, a.ctypes.data, a.size, op, threads) if threads > 1 else self.lib.pgo_count_feed(self.h, a.ctypes.data, a.size, op)
        if st != 0:
            raise RuntimeError(self.lib.pgo_last_error().decode())

    def getKmerAbundance(self, kmer: str) -> int:
        out = np.zeros(1, np.uint64)
        b = np.frombuffer(kmer.encode(), np.uint8)
        self.lib.pgo_count_lookup_ascii(self.h, b.ctypes.data, 1, out.ctypes.data)
        return int(out[0])

    def lookup(self, codes):
        codes = np.ascontiguousarray(codes, np.uint64)
        out = np.zeros(len(codes), np.uint64)
        self.lib.pgo_count_lookup(self.h, codes.ctypes.data, len(codes), out.ctypes.data)
        return out

    def histogram(self, max_count=10000):
        bins = np.zeros(max_count + 1, np.uint64)
        self.lib.pgo_count_histogram(self.h, max_count, bins.ctypes.data)
        return bins

    def computeHistogram(self, max_count, largest_peak, filename=""):
        out = C.c_uint64(0)
        st = self.lib.pgo_count_compute_histogram(self.h, max_count, int(largest_peak), filename.encode() if filename else None, C.byref(out))
        if st != 0:
            raise RuntimeError(self.lib.pgo_last_error().decode())
        return out.value

    def computeKmerCoverage(self, genome_kmers):
        out = C.c_uint64(0)
        self.lib.pgo_count_kmer_coverage(self.h, genome_kmers, C.byref(out))
        return out.value

    def distinct(self):
        return int(self.lib.pgo_count_distinct(self.h))

    def fill_counts(self, peak, panels):
        pa = (PgPanel * len(panels))()
        for i, p in enumerate(panels):
            pa[i] = p.as_struct()
        st = self.lib.pgo_fill_counts(self.h, peak, len(panels), pa)
        if st != 0:
            raise RuntimeError(self.lib.pgo_last_error().decode())

    def __del__(self):
        try:
            self.lib.pgo_count_destroy(self.h)
        except Exception:
            pass


def cpu_haplotype_sample(lib, prefix, panel, size, recombrate=1.26, effective_N=25000.0, add_reference=False, allele_penalty=10):
    """pgo_/pgr_ haplotype_sample -> (sampled_paths [n_out, V], best_scores [size], new_path_to_allele [V, n_out],
    new_kmer_count [V], new_counts) — the reference's HaplotypeSampler / its CPU restatement (oracle/pg_oracle_sampler.cpp)."""
    V, n_out = panel.n_variants, size + (1 if add_reference else 0)
    paths = np.zeros((n_out, V), np.uint64)
    scores = np.zeros(size, np.uint32)
    p2a = np.zeros((V, n_out), np.uint16)
    nk = np.zeros(V, np.uint32)
    counts = np.zeros(max(len(panel.kmer_counts), 1), np.uint16)
    ps = panel.as_struct()
    f = getattr(lib, prefix + "haplotype_sample")
    f.restype = C.c_int
    f.argtypes = [C.POINTER(PgPanel), C.c_uint32, C.c_double, C.c_double, C.c_int, C.c_uint16, C.c_void_p, C.c_void_p,
                  C.c_void_p, C.c_void_p, C.c_void_p]
    st = f(C.byref(ps), size, recombrate, effective_N, int(add_reference), allele_penalty, ptr(paths), ptr(scores), ptr(p2a),
           ptr(nk), ptr(counts))
    if st != 0:
        raise RuntimeError(getattr(lib, prefix + "last_error")().decode())
    return paths, scores, p2a, nk, counts[:int(nk.sum())]


def oracle_unique_kmers(lib, graph_counter: "OracleCounter", flat: dict):
    """pgo_unique_kmers_compute over the flat arrays of tests/refgraph.flatten -> Panel (index stage)."""
    from pangenie_b200.model import panel_from_struct, variants_struct
    vs, _keep = variants_struct(flat)
    h = lib.pgo_unique_kmers_compute(graph_counter.h, C.byref(vs))
    if not h:
        raise RuntimeError("pgo_unique_kmers_compute failed (an undefined allele that no path carries?)")
    try:
        ps = PgPanel()
        assert lib.pgo_unique_kmers_panel(h, C.byref(ps)) == 0
        return panel_from_struct(ps)
    finally:
        lib.pgo_unique_kmers_free(h)
