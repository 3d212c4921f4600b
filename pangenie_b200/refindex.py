"""Readers for the reference's index artefacts, host side only.

* `<prefix>_UniqueKmersMap.cereal` — cereal BinaryOutputArchive of `UniqueKmersMap` (reference src/commands.hpp:12-27):
  raw little-endian; size_t -> u64; string / vector / map = u64 length + elements; vector<bool> one byte per element;
  polymorphic shared_ptr = u32 polymorphic id (MSB set on first use, then u64-length-prefixed class name) + u32 pointer
  id (MSB set = object follows).  Fields of the two concrete classes: reference src/biallelicuniquekmers.hpp:102-114,
  src/multiallelicuniquekmers.hpp:101-113, src/kmerpath.hpp:26-33, src/kmerpath16.hpp:26-33.
* `<prefix>_<chrom>_kmers.tsv.gz` — 5 tab-separated columns, comma lists of k-mers, "nan" if empty
  (reference src/kmerparser.cpp:16-28, header at src/stepwiseuniquekmercomputer.cpp:105).
"""
from __future__ import annotations

import gzip
import struct

import numpy as np

from .panel import Panel

_CODE = {65: 0, 67: 1, 71: 2, 84: 3, 97: 0, 99: 1, 103: 2, 116: 3}


class _Reader:
    def __init__(self, data: bytes):
        self.d, self.o = data, 0

    def take(self, fmt):
        v = struct.unpack_from("<" + fmt, self.d, self.o)
        self.o += struct.calcsize("<" + fmt)
        return v[0] if len(v) == 1 else v

    def string(self):
        n = self.take("Q")
        s = self.d[self.o:self.o + n].decode()
        self.o += n
        return s


def read_unique_kmers_map(path: str):
    """-> (kmersize, {chrom: Panel}, add_reference).  Counts/coverage are whatever the archive holds."""
    r = _Reader(open(path, "rb").read())
    kmersize = r.take("Q")
    names = {}
    chroms = {}
    for _ in range(r.take("Q")):
        chrom = r.string()
        variants = []
        for _v in range(r.take("Q")):
            pid = r.take("I")
            if pid & 0x80000000:
                names[pid & 0x7fffffff] = r.string()
            cls = names[pid & 0x7fffffff]
            ptr = r.take("I")
            assert ptr & 0x80000000, "shared objects are not expected in this archive"
            pos = r.take("Q")
            cov = r.take("f")
            cur = r.take("Q")
            counts = [r.take("H") for _ in range(r.take("Q"))]
            alleles = []
            for _a in range(r.take("Q")):
                if cls == "BiallelicUniqueKmers":
                    aid = r.take("B")
                    off, mask = r.take("H"), r.take("H")
                else:
                    aid = r.take("H")
                    off, mask = r.take("H"), r.take("I")
                undefined = r.take("B")
                alleles.append((aid, off, mask, undefined))
            npth = r.take("Q")
            p2a = [r.take("B") if cls == "BiallelicUniqueKmers" else r.take("H") for _ in range(npth)]
            assert cur == len(counts)
            variants.append((pos, int(cov), counts, alleles, p2a))
        chroms[chrom] = variants
    for _m in range(2):  # runtimes, sampling_runtimes
        for _ in range(r.take("Q")):
            r.string()
            r.take("d")
    add_reference = bool(r.take("B"))
    assert r.o == len(r.d), "trailing bytes in archive"
    panels = {}
    for chrom, variants in chroms.items():
        V = len(variants)
        P = len(variants[0][4]) if V else 0
        koff = np.zeros(V + 1, np.uint32)
        aoff = np.zeros(V + 1, np.uint32)
        kc, aid, aun, ako, akm = [], [], [], [], []
        pos = np.zeros(V, np.uint64)
        cov = np.zeros(V, np.uint16)
        p2a = np.zeros(V * P, np.uint16)
        for v, (ps, cv, counts, alleles, pa) in enumerate(variants):
            pos[v], cov[v] = ps, cv
            p2a[v * P:(v + 1) * P] = pa
            kc += counts
            for a in sorted(alleles):
                aid.append(a[0]); ako.append(a[1]); akm.append(a[2]); aun.append(a[3])
            koff[v + 1], aoff[v + 1] = len(kc), len(aid)
        panels[chrom] = Panel(P, pos, p2a, cov, koff, np.array(kc, np.uint16), aoff, np.array(aid, np.uint16),
                              np.array(aun, np.uint8), np.array(ako, np.uint16), np.array(akm, np.uint32))
    return kmersize, panels, add_reference


def encode_kmer(s: str) -> int:
    v = 0
    for ch in s.encode():
        v = (v << 2) | _CODE[ch]
    return v


def attach_kmers_tsv(panel: Panel, path: str):
    """Fills kmer_codes / flank_offsets / flank_codes of `panel` from `<prefix>_<chrom>_kmers.tsv.gz`."""
    kcodes, fcodes, foff = [], [], [0]
    with gzip.open(path, "rt") as f:
        v = 0
        for line in f:
            t = line.rstrip("\n").split("\t")
            if t[0].startswith("#"):
                continue
            assert int(t[1]) == int(panel.positions[v]), "variant order differs from the UniqueKmersMap"
            ks = [] if t[3] == "nan" else t[3].split(",")
            fs = [] if t[4] == "nan" else t[4].split(",")
            assert len(ks) == int(panel.kmer_offsets[v + 1] - panel.kmer_offsets[v])
            kcodes += [encode_kmer(x) for x in ks]
            fcodes += [encode_kmer(x) for x in fs]
            foff.append(len(fcodes))
            v += 1
    panel.kmer_codes = np.array(kcodes, np.uint64)
    panel.flank_offsets = np.array(foff, np.uint32)
    panel.flank_codes = np.array(fcodes, np.uint64)
    return panel
