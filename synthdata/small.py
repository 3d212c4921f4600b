"""Deterministic synthetic workloads of the shapes BASELINE.json names.

Genome: i.i.d. uniform ACGT.  Variants: SNPs (biallelic, a few tri-allelic), spacing uniform in
[100, 1100] bp so bubbles stay single-record.  Panel: H haplotypes copied Li-Stephens style from 8
founders (+ the reference path 0, so P = H + 1).  Reads: 150 bp, both strands, substitution errors,
4-line FASTQ with fixed-width records.  Index side (what PanGenie-index would have produced): per
variant up to 16 unique k-mers per allele (2-bit codes + allele membership) and up to 24 flanking
k-mers; segment FASTA = reference sequence plus one record per ALT allele with k-1 flanks
(reference src/graphbuilder.cpp:293-353).

Everything here is host-side test/bench tooling (numpy); nothing is on the accelerated path.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from pangenie_b200.panel import Panel

GRCH38_AUTOSOME_MBP = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51]
_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP_ASCII = np.zeros(256, np.uint8)
for _a, _b in zip(b"ACGT", b"TGCA"):
    _COMP_ASCII[_a] = _b


@dataclass
class Chromosome:
    name: str
    genome: np.ndarray        # u8 codes 0..3
    positions: np.ndarray     # i64[V] 0-based SNP positions
    n_alleles: np.ndarray     # u8[V] 2 or 3
    alt: np.ndarray           # u8[V,2] alternative base codes (second used when tri-allelic)
    haplotypes: np.ndarray    # u8[V,H] allele id carried by each panel haplotype
    panel: Panel | None = None


@dataclass
class Workload:
    k: int
    chromosomes: list
    reads_fastq: np.ndarray   # u8 FASTQ text
    segments_fasta: np.ndarray  # u8 FASTA text
    truth: list               # per chromosome (V,2) allele ids of the simulated sample
    n_reads: int
    read_len: int

    @property
    def n_variants(self) -> int:
        return sum(len(c.positions) for c in self.chromosomes)

    @property
    def record_bytes(self) -> int:
        """Bytes of one (fixed-width) FASTQ record: '@' + 9 digits + LF, seq + LF, '+' LF, qual + LF."""
        return 11 + self.read_len + 3 + self.read_len + 1

    @property
    def panels(self):
        return [c.panel for c in self.chromosomes]


def _pack_kmers(windows: np.ndarray, k: int) -> np.ndarray:
    """windows: (..., n) base codes; returns (..., n-k+1) 2-bit k-mer codes (first base most significant)."""
    n = windows.shape[-1]
    out = np.zeros(windows.shape[:-1] + (n - k + 1,), np.uint64)
    for j in range(k):
        out = (out << np.uint64(2)) | windows[..., j:n - k + 1 + j].astype(np.uint64)
    return out


def make_chromosome(rng: np.random.Generator, name: str, n_variants: int, n_haplotypes: int, k: int,
                    tri_frac: float = 0.02) -> Chromosome:
    gaps = rng.integers(100, 1101, size=n_variants)
    positions = np.cumsum(gaps) + 2 * k
    length = int(positions[-1]) + 2 * k + 200 if n_variants else 4 * k + 200
    genome = rng.integers(0, 4, size=length, dtype=np.uint8)
    n_alleles = np.where(rng.random(n_variants) < tri_frac, 3, 2).astype(np.uint8)
    ref = genome[positions]
    alt1 = (ref + rng.integers(1, 4, size=n_variants).astype(np.uint8)) % 4
    # second ALT: one of the two remaining bases
    others = np.array([[b for b in range(4)] for _ in range(1)], np.uint8)
    alt2 = np.zeros(n_variants, np.uint8)
    pick = rng.integers(0, 2, size=n_variants)
    for v_ref in range(4):
        for v_a1 in range(4):
            if v_ref == v_a1:
                continue
            rest = [b for b in range(4) if b not in (v_ref, v_a1)]
            m = (ref == v_ref) & (alt1 == v_a1)
            alt2[m] = np.where(pick[m] == 0, rest[0], rest[1])
    del others
    # founders and Li-Stephens-like copying
    af = np.clip(rng.beta(0.5, 0.5, size=n_variants), 0.02, 0.98)
    founders = (rng.random((n_variants, 8)) < af[:, None]).astype(np.uint8)
    tri = n_alleles == 3
    founders[tri] = np.where(founders[tri] > 0, rng.integers(1, 3, size=(int(tri.sum()), 8)).astype(np.uint8), 0)
    # guarantee a non-reference allele in every column (keeps every variant an HMM column, as SURVEY 8d asks)
    none = founders.max(axis=1) == 0
    founders[none, rng.integers(0, 8, size=int(none.sum()))] = 1
    hap = np.zeros((n_variants, n_haplotypes), np.uint8)
    switch_p = 1.0 - np.exp(-1e-4 * gaps)
    for h in range(n_haplotypes):
        sw = rng.random(n_variants) < switch_p
        sw[0] = True
        src = rng.integers(0, 8, size=n_variants)
        idx = np.maximum.accumulate(np.where(sw, np.arange(n_variants), 0))
        hap[:, h] = founders[np.arange(n_variants), src[idx]]
    # make sure each column still has a non-reference allele on some panel haplotype
    none = hap.max(axis=1) == 0
    hap[none, 0] = 1
    return Chromosome(name, genome, positions.astype(np.int64), n_alleles, np.stack([alt1, alt2], axis=1), hap)


def build_panel(ch: Chromosome, k: int, kmers_per_allele: int = 16, flanks_per_side: int = 12) -> Panel:
    """What PanGenie-index would hand to the genotyper for this chromosome (SNP bubbles)."""
    V, H = ch.haplotypes.shape
    P = H + 1
    pos = ch.positions
    # alleles actually present on paths (path 0 = reference carries allele 0)
    present = np.zeros((V, 3), bool)
    present[:, 0] = True
    for a in (1, 2):
        present[:, a] = (ch.haplotypes == a).any(axis=1)
    # k-mer windows over each allele: offsets -(k-1)..0 relative to the SNP give the k k-mers covering it
    offs = np.arange(-(k - 1), k)
    win = ch.genome[pos[:, None] + offs[None, :]]  # (V, 2k-1)
    sel = np.unique(np.linspace(0, k - 1, kmers_per_allele).round().astype(int))
    n_sel = len(sel)
    codes_by_allele = []
    for a in range(3):
        w = win.copy()
        if a > 0:
            w[:, k - 1] = ch.alt[:, a - 1]
        codes_by_allele.append(_pack_kmers(w, k)[:, sel])  # (V, n_sel)
    n_all = present.sum(axis=1)
    kcount = n_all * n_sel
    koff = np.zeros(V + 1, np.uint32)
    koff[1:] = np.cumsum(kcount)
    aoff = np.zeros(V + 1, np.uint32)
    aoff[1:] = np.cumsum(n_all)
    K, A = int(koff[-1]), int(aoff[-1])
    kcodes = np.zeros(K, np.uint64)
    aid = np.zeros(A, np.uint16)
    akoff = np.zeros(A, np.uint16)
    amask = np.zeros(A, np.uint32)
    rank = np.cumsum(present, axis=1) - 1  # index of allele a within the variant's allele list
    for a in range(3):
        m = present[:, a]
        vi = np.nonzero(m)[0]
        r = rank[vi, a]
        kbase = koff[vi].astype(np.int64) + r * n_sel
        kcodes[(kbase[:, None] + np.arange(n_sel)[None, :]).ravel()] = codes_by_allele[a][vi].ravel()
        ai = aoff[vi].astype(np.int64) + r
        aid[ai] = a
        akoff[ai] = (r * n_sel).astype(np.uint16)
        amask[ai] = (1 << n_sel) - 1
    # flanking k-mers: non-overlapping the SNP, stepping away from it
    fl = []
    for s in range(1, flanks_per_side + 1):
        fl.append(pos - (k - 1) - s * 3 - (k - 1))  # left: window ends before the first covering k-mer starts
        fl.append(pos + 1 + s * 3)                  # right
    fstart = np.stack(fl, axis=1)
    fstart = np.clip(fstart, 0, len(ch.genome) - k)
    fw = ch.genome[fstart[:, :, None] + np.arange(k)[None, None, :]]
    fcodes = _pack_kmers(fw, k)[:, :, 0]
    foff = (np.arange(V + 1) * fcodes.shape[1]).astype(np.uint32)
    p2a = np.zeros((V, P), np.uint16)
    p2a[:, 1:] = ch.haplotypes
    return Panel(P, pos.astype(np.uint64), p2a.ravel(), np.zeros(V, np.uint16), koff, np.zeros(K, np.uint16), aoff, aid,
                 np.zeros(A, np.uint8), akoff, amask, kcodes, foff, fcodes.ravel().astype(np.uint64))


def _fasta_record(name: str, codes: np.ndarray, width: int = 60) -> np.ndarray:
    seq = _ASCII[codes]
    n = len(seq)
    full = (n // width) * width
    body = np.concatenate([seq[:full].reshape(-1, width), np.full((n // width, 1), 10, np.uint8)], axis=1).ravel()
    tail = np.concatenate([seq[full:], np.array([10], np.uint8)]) if n > full else np.zeros(0, np.uint8)
    return np.concatenate([np.frombuffer(f">{name}\n".encode(), np.uint8), body, tail])


def segments_fasta(chroms, k: int) -> np.ndarray:
    parts = []
    for ch in chroms:
        parts.append(_fasta_record(ch.name, ch.genome))
        offs = np.arange(-(k - 1), k)
        for a in (1, 2):
            m = (ch.haplotypes == a).any(axis=1)
            vi = np.nonzero(m)[0]
            if not len(vi):
                continue
            w = ch.genome[ch.positions[vi, None] + offs[None, :]].copy()
            w[:, k - 1] = ch.alt[vi, a - 1]
            # fixed-width records ">s\n" + 2k-1 bases + "\n"
            rec = np.empty((len(vi), 3 + (2 * k - 1) + 1), np.uint8)
            rec[:, 0] = ord(">")
            rec[:, 1] = ord("s")
            rec[:, 2] = 10
            rec[:, 3:3 + 2 * k - 1] = _ASCII[w]
            rec[:, -1] = 10
            parts.append(rec.ravel())
    return np.concatenate(parts)


def simulate_reads(rng: np.random.Generator, chroms, truth, coverage: float, read_len: int = 150, err: float = 0.002):
    """FASTQ text (fixed-width records) from a diploid sample; returns (u8 array, n_reads)."""
    recs = []
    total = 0
    for ch, tr in zip(chroms, truth):
        G = len(ch.genome)
        n = int(coverage * G / read_len)
        if n == 0:
            continue
        haps = []
        for h in range(2):
            s = ch.genome.copy()
            al = tr[:, h]
            m1, m2 = al == 1, al == 2
            s[ch.positions[m1]] = ch.alt[m1, 0]
            s[ch.positions[m2]] = ch.alt[m2, 1]
            haps.append(s)
        start = rng.integers(0, G - read_len + 1, size=n)
        which = rng.integers(0, 2, size=n)
        idx = start[:, None] + np.arange(read_len)[None, :]
        seq = np.where(which[:, None] == 0, haps[0][idx], haps[1][idx]).astype(np.uint8)
        e = rng.random(seq.shape) < err
        seq = np.where(e, (seq + rng.integers(1, 4, size=seq.shape).astype(np.uint8)) % 4, seq).astype(np.uint8)
        asc = _ASCII[seq]
        rc = rng.random(n) < 0.5
        asc[rc] = _COMP_ASCII[asc[rc][:, ::-1]]
        hdr = 11  # "@" + 9 digits + "\n"
        rec = np.empty((n, hdr + read_len + 3 + read_len + 1), np.uint8)
        ids = np.arange(total, total + n)
        rec[:, 0] = ord("@")
        for d in range(9):
            rec[:, 1 + d] = 48 + (ids // 10 ** (8 - d)) % 10
        rec[:, 10] = 10
        rec[:, hdr:hdr + read_len] = asc
        rec[:, hdr + read_len] = 10
        rec[:, hdr + read_len + 1] = ord("+")
        rec[:, hdr + read_len + 2] = 10
        rec[:, hdr + read_len + 3:hdr + 2 * read_len + 3] = ord("F")
        rec[:, -1] = 10
        recs.append(rec.ravel())
        total += n
    return (np.concatenate(recs) if recs else np.zeros(0, np.uint8)), total


def make_workload(config: int | None = None, *, n_chrom: int = 1, n_variants: int = 10_000, n_haplotypes: int = 8,
                  coverage: float = 10.0, k: int = 31, seed: int | None = None, with_reads: bool = True) -> Workload:
    """config 1..3 = BASELINE.json configs[1..3]; otherwise the explicit shape."""
    if config is not None:
        n_chrom, n_variants, n_haplotypes, coverage = {
            1: (1, 10_000, 8, 10.0), 2: (22, 1_000_000, 32, 30.0), 3: (22, 5_000_000, 64, 30.0), 4: (22, 5_000_000, 128, 30.0),
        }[config]
        seed = 20260925 + config if seed is None else seed
    rng = np.random.default_rng(20260925 if seed is None else seed)
    if n_chrom == 22:
        w = np.array(GRCH38_AUTOSOME_MBP, float)
        per = np.maximum(1, np.round(n_variants * w / w.sum()).astype(int))
    else:
        per = np.full(n_chrom, n_variants // n_chrom)
        per[: n_variants - per.sum()] += 1
    chroms = []
    for c in range(n_chrom):
        ch = make_chromosome(rng, f"chr{c + 1}", int(per[c]), n_haplotypes, k)
        ch.panel = build_panel(ch, k)
        chroms.append(ch)
    truth = []
    for ch in chroms:
        H = ch.haplotypes.shape[1]
        V = len(ch.positions)
        t = np.zeros((V, 2), np.uint8)
        gaps = np.diff(ch.positions, prepend=0)
        for h in range(2):  # mosaic of panel haplotypes
            sw = rng.random(V) < (1.0 - np.exp(-2e-5 * gaps))
            if V:
                sw[0] = True
            src = rng.integers(0, H, size=V)
            idx = np.maximum.accumulate(np.where(sw, np.arange(V), 0))
            t[:, h] = ch.haplotypes[np.arange(V), src[idx]]
        truth.append(t)
    if with_reads:
        reads, n_reads = simulate_reads(rng, chroms, truth, coverage)
        segs = segments_fasta(chroms, k)
    else:
        reads, n_reads, segs = np.zeros(0, np.uint8), 0, np.zeros(0, np.uint8)
    return Workload(k, chroms, reads, segs, truth, n_reads, 150)


def fill_synthetic_counts(rng: np.random.Generator, wl: Workload, peak: int = 24):
    """Fills kmer_counts / coverage without counting reads: Poisson(peak * copy number) + error k-mers.
    Used by HMM-only tests and benches."""
    for ch, tr in zip(wl.chromosomes, wl.truth):
        p = ch.panel
        V = p.n_variants
        p.coverage[:] = np.clip(rng.poisson(peak, size=V), peak // 4 + 1, peak * 4 - 1)
        cn_by_allele = np.zeros((V, 3), np.int64)
        for h in range(2):
            np.add.at(cn_by_allele, (np.arange(V), tr[:, h]), 1)
        ao = p.allele_offsets
        for v_a in range(3):
            pass
        # per allele-list entry: copy number, then expand to its k-mers via (offset, mask)
        n_all = np.diff(ao)
        var_of_allele = np.repeat(np.arange(V), n_all)
        cn = cn_by_allele[var_of_allele, p.allele_ids]
        nk = np.array([bin(int(m)).count("1") for m in np.unique(p.allele_kmer_mask)])
        per = int(nk.max()) if len(nk) else 0
        # all alleles carry `per` consecutive k-mers starting at allele_kmer_offset
        kidx = (p.kmer_offsets[var_of_allele].astype(np.int64) + p.allele_kmer_offset)[:, None] + np.arange(per)[None, :]
        lam = np.repeat(cn[:, None], per, axis=1) * (peak / 2.0)
        counts = rng.poisson(lam) + (rng.random(lam.shape) < 0.02) * rng.integers(1, 3, size=lam.shape)
        p.kmer_counts[kidx.ravel()] = np.minimum(counts.ravel(), 65535).astype(np.uint16)
