"""Flat SoA panel (`pg_panel`) and a builder that mirrors the reference's UniqueKmers API.

`PanelBuilder.add_variant / insert_kmer / set_undefined_allele / set_coverage` follow
`BiallelicUniqueKmers` / `MultiallelicUniqueKmers` (reference src/biallelicuniquekmers.cpp:8-48,
src/multiallelicuniquekmers.cpp) so parity tests read like the reference's own tests, e.g.
tests/HMMTest.cpp:14-36.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from .capi import PgHmmResult, PgPanel, ptr


@dataclass
class Panel:
    """One chromosome: `std::vector<std::shared_ptr<UniqueKmers>>` flattened (include/pangenie_b200.h)."""
    n_paths: int
    positions: np.ndarray          # u64[V]
    path_to_allele: np.ndarray     # u16[V*P]
    coverage: np.ndarray           # u16[V]
    kmer_offsets: np.ndarray       # u32[V+1]
    kmer_counts: np.ndarray        # u16[K]
    allele_offsets: np.ndarray     # u32[V+1]
    allele_ids: np.ndarray         # u16[A]
    allele_undefined: np.ndarray   # u8[A]
    allele_kmer_offset: np.ndarray # u16[A]
    allele_kmer_mask: np.ndarray   # u32[A]
    kmer_codes: np.ndarray | None = None     # u64[K]
    flank_offsets: np.ndarray | None = None  # u32[V+1]
    flank_codes: np.ndarray | None = None    # u64[F]

    @property
    def n_variants(self) -> int:
        return int(self.positions.shape[0])

    def as_struct(self) -> PgPanel:
        s = PgPanel()
        s.n_variants = self.n_variants
        s.n_paths = self.n_paths
        for name in ("positions", "path_to_allele", "coverage", "kmer_offsets", "kmer_counts", "allele_offsets",
                     "allele_ids", "allele_undefined", "allele_kmer_offset", "allele_kmer_mask", "kmer_codes",
                     "flank_offsets", "flank_codes"):
            setattr(s, name, ptr(getattr(self, name)))
        return s

    def result_layout(self) -> np.ndarray:
        """gl_offsets as pg_result_layout computes them: row length n(n+1)/2 with n = max allele id + 1."""
        V = self.n_variants
        off = np.zeros(V + 1, dtype=np.uint64)
        if V == 0:
            return off
        ao = self.allele_offsets.astype(np.int64)
        n = np.ones(V, np.int64)
        has = ao[1:] > ao[:-1]
        if has.any():
            mx = np.maximum.reduceat(self.allele_ids.astype(np.int64), ao[:-1][has])
            n[has] = mx + 1
        off[1:] = np.cumsum(n * (n + 1) // 2)
        return off

    def nr_alleles(self, v: int) -> int:
        ids = self.allele_ids[self.allele_offsets[v]:self.allele_offsets[v + 1]]
        return (int(ids.max()) if len(ids) else 0) + 1


class Result:
    """Caller-owned output buffers of one chromosome (`pg_hmm_result`)."""

    def __init__(self, panel: Panel):
        V = panel.n_variants
        self.gl_offsets = panel.result_layout()
        self.likelihoods = np.zeros(int(self.gl_offsets[-1]), dtype=np.float64)
        self.is_column = np.zeros(V, dtype=np.uint8)
        self.genotype = np.zeros(2 * V, dtype=np.int16)
        self.quality = np.zeros(V, dtype=np.uint32)
        self.unique_kmers = np.zeros(V, dtype=np.uint16)
        self.coverage = np.zeros(V, dtype=np.uint16)

    def as_struct(self) -> PgHmmResult:
        s = PgHmmResult()
        for name in ("gl_offsets", "likelihoods", "is_column", "genotype", "quality", "unique_kmers", "coverage"):
            setattr(s, name, ptr(getattr(self, name)))
        return s

    def row(self, v: int) -> np.ndarray:
        return self.likelihoods[int(self.gl_offsets[v]):int(self.gl_offsets[v + 1])]

    def get_genotype_likelihood(self, v: int, a1: int, a2: int) -> float:
        """GenotypingResult::get_genotype_likelihood (reference src/genotypingresult.cpp:39-46)."""
        lo, hi = (a1, a2) if a1 <= a2 else (a2, a1)
        idx = hi * (hi + 1) // 2 + lo
        r = self.row(v)
        return float(r[idx]) if idx < len(r) else 0.0


@dataclass
class _Variant:
    position: int
    path_to_allele: list
    counts: list = field(default_factory=list)
    on_alleles: list = field(default_factory=list)   # per k-mer: list of allele ids
    undefined: set = field(default_factory=set)
    coverage: int = 0
    codes: list = field(default_factory=list)
    flanks: list = field(default_factory=list)


class PanelBuilder:
    def __init__(self):
        self._vars: list[_Variant] = []

    def add_variant(self, position: int, path_to_allele) -> int:
        """`BiallelicUniqueKmers(position, alleles)` / `MultiallelicUniqueKmers(position, alleles)`."""
        self._vars.append(_Variant(int(position), [int(a) for a in path_to_allele]))
        return len(self._vars) - 1

    def insert_kmer(self, v: int, readcount: int, alleles, code: int = 0):
        """`insert_kmer(readcount, allele_ids)` (src/biallelicuniquekmers.cpp:38-48)."""
        var = self._vars[v]
        var.counts.append(int(readcount))
        var.on_alleles.append([int(a) for a in alleles])
        var.codes.append(int(code))

    def set_undefined_allele(self, v: int, allele: int):
        self._vars[v].undefined.add(int(allele))

    def set_coverage(self, v: int, cov: int):
        self._vars[v].coverage = int(cov)

    def set_flanks(self, v: int, codes):
        self._vars[v].flanks = [int(c) for c in codes]

    def build(self, with_codes: bool = False) -> Panel:
        V = len(self._vars)
        P = len(self._vars[0].path_to_allele) if V else 0
        pos = np.zeros(V, np.uint64)
        p2a = np.zeros(V * P, np.uint16)
        cov = np.zeros(V, np.uint16)
        koff = np.zeros(V + 1, np.uint32)
        aoff = np.zeros(V + 1, np.uint32)
        foff = np.zeros(V + 1, np.uint32)
        kcnt, aid, aund, akoff, amask, kcodes, fcodes = [], [], [], [], [], [], []
        for v, var in enumerate(self._vars):
            assert len(var.path_to_allele) == P, "all variants must be covered by the same paths"
            pos[v] = var.position
            p2a[v * P:(v + 1) * P] = var.path_to_allele
            cov[v] = var.coverage
            kcnt += var.counts
            kcodes += var.codes
            fcodes += var.flanks
            koff[v + 1] = len(kcnt)
            foff[v + 1] = len(fcodes)
            for a in sorted(set(var.path_to_allele)):   # the `alleles` map is keyed by the alleles on paths
                ks = [k for k, on in enumerate(var.on_alleles) if a in on]
                off = ks[0] if ks else 0
                mask = 0
                for k in ks:
                    if k - off >= 32:
                        raise ValueError("KmerPath::KmerPath: index is invalid")  # src/kmerpath.cpp:23-26
                    mask |= 1 << (k - off)
                aid.append(a)
                aund.append(1 if a in var.undefined else 0)
                akoff.append(off)
                amask.append(mask)
            aoff[v + 1] = len(aid)
        return Panel(P, pos, p2a, cov, koff, np.array(kcnt, np.uint16), aoff, np.array(aid, np.uint16),
                     np.array(aund, np.uint8), np.array(akoff, np.uint16), np.array(amask, np.uint32),
                     np.array(kcodes, np.uint64) if with_codes else None,
                     foff if with_codes else None,
                     np.array(fcodes, np.uint64) if with_codes else None)
