"""Synthetic PanGenie workloads at the full BASELINE.json sizes, generated with torch on the GPU
when one is present (a 30x / 600 Mbp sample is 38 GB of FASTQ: a few seconds on the device, minutes in numpy).

Test and bench tooling only.  What is generated, per chromosome `chrNN` (all random streams are seeded per chromosome and
per read chunk, so any rank can generate any part of the sample independently and every rank sees the same sample):

  genome     i.i.d. uniform ACGT
  variants   spacing uniform in [100, 1100] bp; 90 % biallelic SNPs, 8 % biallelic indels (1-50 bp, VCF-style anchored),
             2 % tri-allelic SNPs; 1 % of the variants additionally carry an UNDEFINED allele (no sequence, no k-mers) on a
             few panel haplotypes
  panel      H haplotypes copied Li-Stephens style from 8 founders (switch rate 1e-4 / bp, allele frequencies
             Beta(0.5, 0.5) clipped to [0.02, 0.98]) + the reference path 0  (P = H + 1, reference pangenie-index.cpp:25)
  index      what PanGenie-index hands to the genotyper (reference src/stepwiseuniquekmercomputer.cpp:46-93, 99-195,
             227-264): per variant the k-mers that occur on exactly one allele sequence (k-1 flanks included), at most
             16 (biallelic) / 32 per allele in ascending k-mer order, allele-major; up to 12 + 12 flanking k-mers from the 2k
             overhangs; the segment FASTA in the layout of GraphBuilder::write_path_segments (src/graphbuilder.cpp:293-352:
             one single-line record per inter-variant reference stretch and one per allele sequence)
  reads      150 bp, both strands, 0.2 % substitutions, uniform starts on a diploid mosaic of two panel haplotypes with
             0.1 % private mismatches; 4-line FASTQ, fixed-width records, file order = chromosome by chromosome
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, replace

import numpy as np
import torch

from pangenie_b200.panel import Panel

GRCH38_AUTOSOME_MBP = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51]
READ_CHUNK = 1 << 20   # reads per independently seeded chunk
HDR_R = 28             # ">chrNN_reference_##########\n"
HDR_A = 20             # ">chrNN_##########_#\n"


@dataclass(frozen=True)
class Spec:
    n_chrom: int
    n_variants: int
    n_haplotypes: int
    coverage: float
    k: int = 31
    seed: int = 20260925
    read_len: int = 150
    err: float = 0.002
    frac_indel: float = 0.08
    frac_tri: float = 0.02
    frac_undef: float = 0.01
    max_indel: int = 50
    text: str = ""


# BASELINE.json configs[1..4]; seed = 20260925 + config index
CONFIGS = {
    "cfg2": Spec(1, 10_000, 8, 10.0, seed=20260926, text="synthetic 1 chrom, 10k variants, 8 haplotypes, 10x reads, k=31 (BASELINE.json configs[1])"),
    "cfg3": Spec(22, 1_000_000, 32, 30.0, seed=20260927, text="synthetic 22 chroms, 1M variants, 32 haplotypes, 30x reads, k=31 (BASELINE.json configs[2])"),
    "cfg4": Spec(22, 5_000_000, 64, 30.0, seed=20260928, text="synthetic 22 chroms, 5M variants, 64 haplotypes, 30x reads, k=31 (BASELINE.json configs[3])"),
    "cfg5": Spec(22, 5_000_000, 128, 30.0, seed=20260929, text="synthetic 22 chroms, 5M variants, 128 haplotypes, 30x reads, k=31, -a 129 (BASELINE.json configs[4])"),
    # reduced shapes for tests and quick runs (same per-column shape, fewer variants)
    "cfg3s": Spec(22, 100_000, 32, 30.0, seed=20260927, text="synthetic 22 chroms, 100k variants, 32 haplotypes, 30x reads, k=31 (configs[2] shape at 1/10 of the variants)"),
    "cfg4s": Spec(22, 50_000, 64, 30.0, seed=20260928, text="synthetic 22 chroms, 50k variants, 64 haplotypes, 30x reads, k=31 (configs[3] shape at 1/100 of the variants)"),
    "tiny": Spec(3, 600, 8, 8.0, seed=20260930, text="synthetic 3 chroms, 600 variants, 8 haplotypes, 8x reads, k=31 (contract tests)"),
}


def variants_per_chrom(spec: Spec) -> list[int]:
    if spec.n_chrom == 22:
        w = np.array(GRCH38_AUTOSOME_MBP, float)
        return [int(x) for x in np.maximum(1, np.round(spec.n_variants * w / w.sum()).astype(int))]
    per = np.full(spec.n_chrom, spec.n_variants // spec.n_chrom)
    per[: spec.n_variants - per.sum()] += 1
    return [int(x) for x in per]


def _gen(device, *key) -> torch.Generator:
    g = torch.Generator(device=device)
    s = 0x9E3779B97F4A7C15
    for x in key:
        s = (s * 6364136223846793005 + int(x) + 1442695040888963407) & 0x7FFFFFFFFFFFFFFF
    g.manual_seed(s)
    return g


@dataclass
class Chrom:
    index: int
    name: str
    genome: torch.Tensor     # u8 [G] codes 0..3
    pos: torch.Tensor        # i64 [V] 0-based start of the REF allele
    ref_len: torch.Tensor    # i64 [V]
    n_alt: torch.Tensor      # i64 [V] 1 or 2
    alt_len: torch.Tensor    # i64 [V, 2]
    alt_seq: torch.Tensor    # u8  [V, 2, max_indel + 1]
    undef: torch.Tensor      # bool [V] variant carries an undefined allele (id n_alt + 1) on some haplotypes
    hap: torch.Tensor        # u8 [V, H] allele id per panel haplotype
    truth: torch.Tensor      # u8 [V, 2] allele ids of the simulated diploid sample
    n_reads: int = 0
    panel: Panel | None = None

    @property
    def V(self) -> int:
        return int(self.pos.shape[0])


def make_chrom(spec: Spec, c: int, device) -> Chrom:
    g = _gen(device, spec.seed, c, 1)
    k, MA = spec.k, spec.max_indel + 1
    V = variants_per_chrom(spec)[c]
    H = spec.n_haplotypes
    kw = dict(generator=g, device=device)
    gaps = torch.randint(100, 1101, (V,), **kw)
    pos = torch.cumsum(gaps, 0) + 2 * k
    G = int(pos[-1].item()) + 2 * k + 200 + spec.max_indel if V else 4 * k + 200
    genome = torch.randint(0, 4, (G,), dtype=torch.uint8, **kw)
    u = torch.rand(V, **kw)
    is_tri = u < spec.frac_tri
    is_indel = (u >= spec.frac_tri) & (u < spec.frac_tri + spec.frac_indel)
    is_del = is_indel & (torch.rand(V, **kw) < 0.5)
    is_ins = is_indel & ~is_del
    ilen = torch.randint(1, spec.max_indel + 1, (V,), **kw)
    ref_len = torch.where(is_del, 1 + ilen, torch.ones_like(ilen))
    ref_base = genome[pos].to(torch.int64)
    d1 = torch.randint(1, 4, (V,), **kw)
    d2 = 1 + (d1 - 1 + torch.randint(1, 3, (V,), **kw)) % 3
    alt_seq = torch.randint(0, 4, (V, 2, MA), dtype=torch.uint8, **kw)   # random inserted bases; only [:alt_len] is used
    alt_seq[:, 0, 0] = torch.where(is_indel, ref_base, (ref_base + d1) % 4).to(torch.uint8)
    alt_seq[:, 1, 0] = ((ref_base + d2) % 4).to(torch.uint8)
    alt_len = torch.ones((V, 2), dtype=torch.int64, device=device)
    alt_len[:, 0] = torch.where(is_ins, 1 + ilen, alt_len[:, 0])
    n_alt = torch.where(is_tri, 2, 1).to(torch.int64)
    # panel haplotypes: founders + Li-Stephens-like copying
    af = torch.sin(torch.rand(V, **kw) * (math.pi / 2)) ** 2          # Beta(1/2, 1/2) is the arcsine law
    af = af.clamp(0.02, 0.98)
    founders = (torch.rand(V, 8, **kw) < af[:, None]).to(torch.uint8)
    second = torch.randint(1, 3, (V, 8), dtype=torch.uint8, **kw)
    founders = torch.where(is_tri[:, None] & (founders > 0), second, founders)
    none = founders.max(dim=1).values == 0
    pick = torch.randint(0, 8, (V,), **kw)
    founders[none, pick[none]] = 1
    sw = torch.rand(V, H, **kw) < (1.0 - torch.exp(-1e-4 * gaps.to(torch.float64)))[:, None]
    if V:
        sw[0] = True
    ar = torch.arange(V, device=device)[:, None].expand(V, H)
    idx = torch.cummax(torch.where(sw, ar, torch.zeros_like(ar)), dim=0).values
    src = torch.randint(0, 8, (V, H), **kw).gather(0, idx)
    hap = founders.gather(1, src)
    none = hap.max(dim=1).values == 0
    hap[none, 0] = 1
    # the simulated sample: a mosaic of two panel haplotypes (never the undefined allele)
    truth = torch.zeros((V, 2), dtype=torch.uint8, device=device)
    for h in range(2):
        sw1 = torch.rand(V, **kw) < (1.0 - torch.exp(-2e-5 * gaps.to(torch.float64)))
        if V:
            sw1[0] = True
        i1 = torch.cummax(torch.where(sw1, ar[:, 0], torch.zeros_like(ar[:, 0])), dim=0).values
        s1 = torch.randint(0, H, (V,), **kw)[i1]
        truth[:, h] = hap.gather(1, s1[:, None])[:, 0]
    undef = torch.rand(V, **kw) < spec.frac_undef
    carriers = (torch.rand(V, H, **kw) < 0.06) & undef[:, None]
    hap = torch.where(carriers, (n_alt + 1).to(torch.uint8)[:, None], hap)
    undef = carriers.any(dim=1)
    G_eff = genome.numel()
    ch = Chrom(c, f"chr{c + 1:02d}", genome, pos, ref_len, n_alt, alt_len, alt_seq, undef, hap, truth)
    ch.n_reads = int(spec.coverage * G_eff / spec.read_len)
    return ch


# ------------------------------------------------------------------------------------------------------
# ragged copy: dst[dst_start[i] + j] = src[src_start[i] + j] for j < lens[i]
# ------------------------------------------------------------------------------------------------------
def ragged_copy(dst, dst_start, src, src_start, lens, lut=None, max_elems: int = 1 << 26):
    n = int(lens.numel())
    if n == 0:
        return
    cs = torch.cumsum(lens, 0)
    total = int(cs[-1].item())
    if total == 0:
        return
    # pieces are processed in runs of at most max_elems elements (temporaries are 8-byte index arrays)
    bounds = [0]
    if total > max_elems:
        targets = torch.arange(max_elems, total, max_elems, device=lens.device)
        bounds += [int(x) for x in torch.searchsorted(cs, targets, right=False).tolist()]
    bounds.append(n)
    for a, b in zip(bounds[:-1], bounds[1:]):
        if b <= a:
            continue
        ln = lens[a:b]
        c0 = torch.cumsum(ln, 0) - ln
        tot = int(ln.sum().item())
        if tot == 0:
            continue
        piece = torch.repeat_interleave(torch.arange(b - a, device=lens.device), ln, output_size=tot)
        within = torch.arange(tot, device=lens.device) - c0[piece]
        vals = src[src_start[a:b][piece] + within]
        if lut is not None:
            vals = lut[vals.to(torch.int64)]
        dst[dst_start[a:b][piece] + within] = vals


def _digits(x: torch.Tensor, n: int) -> torch.Tensor:
    """[len(x), n] ASCII digits of x, zero padded."""
    p = torch.tensor([10 ** (n - 1 - i) for i in range(n)], dtype=torch.int64, device=x.device)
    return ((x[:, None] // p[None, :]) % 10 + 48).to(torch.uint8)


def _ascii_lut(device):
    return torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)


def allele_sequences(spec: Spec, ch: Chrom):
    """Allele sequences with their k-1 flanks (Variant::get_allele_sequence): (S u8 [V, 3, MAXS], len i64 [V, 3])."""
    k, MA = spec.k, spec.max_indel + 1
    V, dev = ch.V, ch.pos.device
    MAXS = 2 * (k - 1) + MA
    s = torch.arange(MAXS, device=dev)[None, :]
    alen = torch.stack([ch.ref_len, ch.alt_len[:, 0], ch.alt_len[:, 1]], dim=1)           # [V, 3]
    S = torch.zeros((V, 3, MAXS), dtype=torch.uint8, device=dev)
    G = ch.genome.numel()
    for a in range(3):
        al = alen[:, a:a + 1]
        left = (ch.pos[:, None] - (k - 1) + s).clamp(0, G - 1)
        right = (ch.pos[:, None] + ch.ref_len[:, None] + (s - (k - 1) - al)).clamp(0, G - 1)
        mid_i = (s - (k - 1)).clamp(0, MA - 1).expand(V, MAXS)
        if a == 0:
            mid = ch.genome[(ch.pos[:, None] + (s - (k - 1)).clamp(0, MA - 1)).clamp(0, G - 1)]
        else:
            mid = ch.alt_seq[:, a - 1, :].gather(1, mid_i)
        base = torch.where(s < k - 1, ch.genome[left], torch.where(s < k - 1 + al, mid, ch.genome[right]))
        S[:, a, :] = base
    return S, alen + 2 * (k - 1)


def _window_codes(S: torch.Tensor, k: int) -> torch.Tensor:
    """[..., n] base codes -> [..., n-k+1] 2-bit k-mer codes (first base most significant), int64 (k <= 31)."""
    n = S.shape[-1]
    W = n - k + 1
    out = torch.zeros(S.shape[:-1] + (W,), dtype=torch.int64, device=S.device)
    for j in range(k):
        out = (out << 2) | S[..., j:j + W].to(torch.int64)
    return out


def build_panel(spec: Spec, ch: Chrom) -> Panel:
    """The index side for one chromosome (see the module docstring); returns host arrays."""
    assert spec.k <= 31
    k, dev, V, H = spec.k, ch.pos.device, ch.V, spec.n_haplotypes
    P = H + 1
    BIG = torch.iinfo(torch.int64).max
    S, slen = allele_sequences(spec, ch)
    W = _window_codes(S, k)                                   # [V, 3, MAXW]
    MAXW = W.shape[-1]
    w = torch.arange(MAXW, device=dev)[None, None, :]
    nwin = slen - k + 1                                       # [V, 3]
    hap = ch.hap.to(torch.int64)
    present = torch.stack([torch.ones(V, dtype=torch.bool, device=dev), (hap == 1).any(1), (hap == 2).any(1) & (ch.n_alt == 2)], dim=1)
    # the sequence of an allele nobody carries is still part of the variant record (it was in the VCF)
    exists = torch.stack([torch.ones(V, dtype=torch.bool, device=dev), torch.ones(V, dtype=torch.bool, device=dev), ch.n_alt == 2], dim=1)
    valid = (w < nwin[:, :, None]) & exists[:, :, None]
    # a k-mer qualifies if it occurs exactly once over all allele sequences of the variant (stepwise_unique_kmers +
    # `local_count > 1` rule, stepwiseuniquekmercomputer.cpp:11-35, 60-61) and its allele is covered by a path (:64-67)
    flat = torch.where(valid, W, -1 - torch.arange(3 * MAXW, device=dev).view(1, 3, MAXW)).reshape(V, 3 * MAXW)
    srt, order = torch.sort(flat, dim=1)
    dup = torch.zeros_like(srt, dtype=torch.bool)
    eq = srt[:, 1:] == srt[:, :-1]
    dup[:, 1:] |= eq
    dup[:, :-1] |= eq
    uniq = torch.zeros_like(dup)
    uniq.scatter_(1, order, ~dup)
    cand = valid & uniq.view(V, 3, MAXW) & present[:, :, None]
    biallelic = (hap <= 1).all(dim=1)
    max_k = torch.where(biallelic, 16, 32)                                               # :75-76
    csort = torch.sort(torch.where(cand, W, torch.full_like(W, BIG)), dim=2).values[:, :, :32]   # ascending k-mer order (std::map<mer_dna>)
    n_sel = torch.minimum(cand.sum(dim=2), max_k[:, None])                               # [V, 3]
    take = torch.arange(32, device=dev)[None, None, :] < n_sel[:, :, None]
    kcodes = csort[take]                                                                 # (variant, allele, rank) order
    koff = torch.zeros(V + 1, dtype=torch.int64, device=dev)
    koff[1:] = torch.cumsum(n_sel.sum(dim=1), 0)
    # alleles map: ids carried by some path, ascending; the undefined allele has id n_alt + 1
    und_id = ch.n_alt + 1
    a_present = torch.cat([present, ch.undef[:, None]], dim=1)                           # [V, 4]
    a_ids = torch.stack([torch.zeros_like(und_id), torch.ones_like(und_id), torch.full_like(und_id, 2), und_id], dim=1)
    a_koff = torch.cat([torch.cumsum(n_sel, 1) - n_sel, torch.zeros((V, 1), dtype=torch.int64, device=dev)], dim=1)
    a_n = torch.cat([n_sel, torch.zeros((V, 1), dtype=torch.int64, device=dev)], dim=1)
    a_koff = torch.where(a_n > 0, a_koff, torch.zeros_like(a_koff))
    a_mask = torch.where(a_n >= 32, torch.full_like(a_n, 0xFFFFFFFF), (torch.ones_like(a_n) << a_n) - 1)
    a_und = torch.zeros((V, 4), dtype=torch.uint8, device=dev)
    a_und[:, 3] = 1
    aoff = torch.zeros(V + 1, dtype=torch.int64, device=dev)
    aoff[1:] = torch.cumsum(a_present.sum(dim=1), 0)
    # flanking k-mers (determine_unique_flanking_kmers, :227-264): 2k overhangs left and right, k-mers in ascending order,
    # at most 12 per side; windows reaching into a neighbouring variant's allele records are not unique in the graph
    G = ch.genome.numel()
    end = ch.pos + ch.ref_len
    prev_end = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), end[:-1]])
    next_pos = torch.cat([ch.pos[1:], torch.full((1,), G, dtype=torch.int64, device=dev)])
    i = torch.arange(k + 1, device=dev)[None, :]
    fl = []
    for side in range(2):
        start = (ch.pos[:, None] - 2 * k + i) if side == 0 else (end[:, None] + i)
        ok = (start >= prev_end[:, None] + (k - 1)) & (start + k <= next_pos[:, None] - (k - 1)) & (start >= 0) & (start + k <= G)
        seq = ch.genome[(start[:, :, None] + torch.arange(k, device=dev)[None, None, :]).clamp(0, G - 1)]
        codes = _window_codes(seq, k)[:, :, 0]
        cs_ = torch.sort(torch.where(ok, codes, torch.full_like(codes, BIG)), dim=1).values[:, :12]
        fl.append((cs_, torch.minimum(ok.sum(dim=1), torch.full_like(ch.pos, 12))))
    ftake = torch.cat([torch.arange(12, device=dev)[None, :] < fl[0][1][:, None], torch.arange(12, device=dev)[None, :] < fl[1][1][:, None]], dim=1)
    fcodes = torch.cat([fl[0][0], fl[1][0]], dim=1)[ftake]
    foff = torch.zeros(V + 1, dtype=torch.int64, device=dev)
    foff[1:] = torch.cumsum(fl[0][1] + fl[1][1], 0)
    p2a = torch.zeros((V, P), dtype=torch.int16, device=dev)
    p2a[:, 1:] = ch.hap.to(torch.int16)

    def h(t, dt):
        return np.ascontiguousarray(t.cpu().numpy().astype(dt, copy=False))
    K, A = int(koff[-1].item()), int(aoff[-1].item())
    return Panel(P, h(ch.pos, np.uint64), h(p2a.reshape(-1), np.uint16), np.zeros(V, np.uint16), h(koff, np.uint32), np.zeros(K, np.uint16),
                 h(aoff, np.uint32), h(a_ids[a_present], np.uint16), h(a_und[a_present], np.uint8), h(a_koff[a_present], np.uint16),
                 h(a_mask[a_present], np.uint32), h(kcodes, np.uint64), h(foff, np.uint32), h(fcodes, np.uint64))


def segments_text(spec: Spec, ch: Chrom):
    """<prefix>_path_segments.fasta for one chromosome (GraphBuilder::write_path_segments, src/graphbuilder.cpp:293-352).
    Returns (text, number of k-mer windows in it = an upper bound on its distinct k-mers)."""
    k, dev, V = spec.k, ch.pos.device, ch.V
    lut = _ascii_lut(dev)
    S, slen = allele_sequences(spec, ch)
    MAXS = S.shape[-1]
    end = ch.pos + ch.ref_len
    prev_end = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), end[:-1]])
    ulen = ch.pos - prev_end
    exists = torch.stack([torch.ones(V, dtype=torch.bool, device=dev), torch.ones(V, dtype=torch.bool, device=dev), ch.n_alt == 2], dim=1)
    rec_len = torch.where(exists, HDR_A + slen + 1, torch.zeros_like(slen))              # [V, 3]
    blk_len = HDR_R + ulen + 1 + rec_len.sum(dim=1)
    blk_start = torch.cumsum(blk_len, 0) - blk_len
    G = ch.genome.numel()
    tail_hdr = f">{ch.name}_reference_end\n".encode()
    last_end = int(end[-1].item()) if V else 0
    total = (int(blk_len.sum().item()) if V else 0) + len(tail_hdr) + (G - last_end) + 1
    out = torch.empty(total, dtype=torch.uint8, device=dev)
    name = torch.tensor(list(ch.name.encode()), dtype=torch.uint8, device=dev)
    if V:
        hr = torch.empty((V, HDR_R), dtype=torch.uint8, device=dev)
        hr[:, 0] = ord(">")
        hr[:, 1:6] = name
        hr[:, 6:17] = torch.tensor(list(b"_reference_"), dtype=torch.uint8, device=dev)
        hr[:, 17:27] = _digits(ch.pos, 10)
        hr[:, 27] = 10
        out[(blk_start[:, None] + torch.arange(HDR_R, device=dev)[None, :]).reshape(-1)] = hr.reshape(-1)
        ragged_copy(out, blk_start + HDR_R, ch.genome, prev_end, ulen, lut=lut)
        out[blk_start + HDR_R + ulen] = 10
        rec_start = blk_start[:, None] + HDR_R + ulen[:, None] + 1 + torch.cumsum(rec_len, 1) - rec_len
        Sflat = S.reshape(-1)
        for a in range(3):
            m = exists[:, a]
            nv = int(m.sum().item())
            if nv == 0:
                continue
            ha = torch.empty((nv, HDR_A), dtype=torch.uint8, device=dev)
            ha[:, 0] = ord(">")
            ha[:, 1:6] = name
            ha[:, 6] = ord("_")
            ha[:, 7:17] = _digits(ch.pos[m], 10)
            ha[:, 17] = ord("_")
            ha[:, 18] = 48 + a
            ha[:, 19] = 10
            rs = rec_start[m, a]
            out[(rs[:, None] + torch.arange(HDR_A, device=dev)[None, :]).reshape(-1)] = ha.reshape(-1)
            vi = torch.nonzero(m)[:, 0]
            ragged_copy(out, rs + HDR_A, Sflat, (vi * 3 + a) * MAXS, slen[m, a], lut=lut)
            out[rs + HDR_A + slen[m, a]] = 10
    t0 = total - (len(tail_hdr) + (G - last_end) + 1)
    out[t0:t0 + len(tail_hdr)] = torch.tensor(list(tail_hdr), dtype=torch.uint8, device=dev)
    out[t0 + len(tail_hdr):total - 1] = lut[ch.genome[last_end:].to(torch.int64)]
    out[total - 1] = 10
    windows = max(0, G - last_end - k + 1)
    if V:
        windows += int((ulen - k + 1).clamp(min=0).sum().item()) + int(torch.where(exists, slen - k + 1, torch.zeros_like(slen)).sum().item())
    return out, windows


def haplotype_sequence(spec: Spec, ch: Chrom, alleles: torch.Tensor, g: torch.Generator) -> torch.Tensor:
    """Base codes of one haplotype carrying `alleles` [V] (+ 0.1 % private substitutions)."""
    dev, V = ch.pos.device, ch.V
    G = ch.genome.numel()
    al = alleles.to(torch.int64)
    end = ch.pos + ch.ref_len
    prev_end = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), end[:-1]])
    ulen = ch.pos - prev_end
    is_alt = al > 0
    alen = torch.where(is_alt, ch.alt_len.gather(1, (al - 1).clamp(0, 1)[:, None])[:, 0], ch.ref_len)
    blk = ulen + alen
    bstart = torch.cumsum(blk, 0) - blk
    last_end = int(end[-1].item()) if V else 0
    total = (int(blk.sum().item()) if V else 0) + (G - last_end)
    out = torch.empty(total, dtype=torch.uint8, device=dev)
    if V:
        ragged_copy(out, bstart, ch.genome, prev_end, ulen)
        r = ~is_alt
        ragged_copy(out, (bstart + ulen)[r], ch.genome, ch.pos[r], alen[r])
        MA = ch.alt_seq.shape[-1]
        vi = torch.nonzero(is_alt)[:, 0]
        ragged_copy(out, (bstart + ulen)[is_alt], ch.alt_seq.reshape(-1), (vi * 2 + (al[is_alt] - 1)) * MA, alen[is_alt])
    out[total - (G - last_end):] = ch.genome[last_end:]
    mm = torch.rand(total, generator=g, device=dev) < 0.001
    out = torch.where(mm, (out + torch.randint(1, 4, (total,), dtype=torch.uint8, generator=g, device=dev)) % 4, out)
    return out


def record_bytes(spec: Spec) -> int:
    """'@' + 9 digits + LF, seq + LF, '+' LF, qual + LF."""
    return 11 + spec.read_len + 3 + spec.read_len + 1


def reads_chunk(spec: Spec, ch: Chrom, haps: torch.Tensor, hap_off, hap_len, chunk: int, first_id: int, n: int) -> torch.Tensor:
    """FASTQ text [n * record_bytes] of reads [chunk * READ_CHUNK, +n) of this chromosome."""
    dev, L = ch.pos.device, spec.read_len
    g = _gen(dev, spec.seed, ch.index, 7, chunk)
    kw = dict(generator=g, device=dev)
    which = torch.randint(0, 2, (n,), **kw)
    hl = torch.where(which == 0, hap_len[0], hap_len[1])
    start = (torch.rand(n, dtype=torch.float64, **kw) * (hl - L + 1).to(torch.float64)).to(torch.int64)
    base = torch.where(which == 0, hap_off[0], hap_off[1]) + start
    rc = torch.rand(n, **kw) < 0.5
    j = torch.arange(L, device=dev)[None, :]
    idx = base[:, None] + torch.where(rc[:, None], L - 1 - j, j)
    seq = haps[idx]
    seq = torch.where(rc[:, None], 3 - seq, seq)
    e = torch.rand((n, L), **kw) < spec.err
    seq = torch.where(e, (seq + torch.randint(1, 4, (n, L), dtype=torch.uint8, **kw)) % 4, seq)
    RB = record_bytes(spec)
    rec = torch.empty((n, RB), dtype=torch.uint8, device=dev)
    rec[:, 0] = ord("@")
    rec[:, 1:10] = _digits(first_id + torch.arange(n, device=dev), 9)
    rec[:, 10] = 10
    rec[:, 11:11 + L] = _ascii_lut(dev)[seq.to(torch.int64)]
    rec[:, 11 + L] = 10
    rec[:, 12 + L] = ord("+")
    rec[:, 13 + L] = 10
    rec[:, 14 + L:14 + 2 * L] = ord("F")
    rec[:, RB - 1] = 10
    return rec.reshape(-1)


@dataclass
class Workload:
    spec: Spec
    device: object
    chrom_variants: list             # variants per chromosome (all chromosomes of the sample)
    chrom_reads: list                # reads per chromosome
    my_chroms: list                  # chromosome indices whose panels this rank holds
    panels: list                     # Panel per entry of my_chroms (host arrays)
    segments: torch.Tensor | None    # u8 FASTA text of ALL chromosomes (device)
    reads: torch.Tensor | None       # u8 FASTQ text of this rank's record range (device)
    read_range: tuple = (0, 0)       # records [a, b) of the sample held in `reads`
    truth: list = field(default_factory=list)
    segment_offsets: list = field(default_factory=list)   # byte offset of every chromosome's records in `segments` (+ total)
    segment_windows: int = 0         # k-mer windows in `segments`: upper bound on the distinct graph k-mers

    @property
    def k(self) -> int:
        return self.spec.k

    @property
    def n_variants(self) -> int:
        return int(sum(self.chrom_variants))

    @property
    def n_reads(self) -> int:
        return int(sum(self.chrom_reads))

    @property
    def record_bytes(self) -> int:
        return record_bytes(self.spec)


def chrom_plan(spec: Spec, device) -> tuple[list[int], list[int]]:
    """(genome length, number of reads) of every chromosome without generating it (the gap stream is the first draw)."""
    lens, reads = [], []
    for c, V in enumerate(variants_per_chrom(spec)):
        g = _gen(device, spec.seed, c, 1)
        gaps = torch.randint(100, 1101, (V,), generator=g, device=device)
        G = int(gaps.sum().item()) + 2 * spec.k + 2 * spec.k + 200 + spec.max_indel if V else 4 * spec.k + 200
        lens.append(G)
        reads.append(int(spec.coverage * G / spec.read_len))
    return lens, reads


def segments_bytes_estimate(spec: Spec, device) -> int:
    """Expected size of the whole segment FASTA without generating it (exact genome lengths, expected allele lengths)."""
    lens, _ = chrom_plan(spec, device)
    V, k = spec.n_variants, spec.k
    e_indel = (1 + spec.max_indel) / 2.0
    e_ref = 1 + spec.frac_indel * 0.5 * e_indel
    e_alt = 1 + spec.frac_indel * 0.5 * e_indel
    per_variant = HDR_R + 1 + 2 * (HDR_A + 1 + 2 * (k - 1)) + e_alt + spec.frac_tri * (HDR_A + 1 + 2 * (k - 1) + 1)
    return int(sum(lens) + V * per_variant + spec.n_chrom * (len(">chr00_reference_end\n") + 1))


def make_workload(spec: Spec, device=None, *, chroms=None, read_records=None, with_segments=True, with_reads=True,
                  with_panels=True, reads_out: torch.Tensor | None = None, segment_chroms=None) -> Workload:
    """Generates (a part of) the sample.

    chroms          chromosome indices whose panels are built (default: all)
    read_records    (a, b): only records [a, b) of the sample's FASTQ are generated (default: all)
    reads_out       optional preallocated u8 tensor to write the FASTQ into
    segment_chroms  chromosomes whose segment records are generated (default: all - the genotyper PRIMEs the whole graph)
    """
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    device = torch.device(device)
    per = variants_per_chrom(spec)
    my = list(range(spec.n_chrom)) if chroms is None else sorted(chroms)
    RB = record_bytes(spec)
    _lens, n_reads = chrom_plan(spec, device)
    total_reads = int(sum(n_reads))
    a, b = (0, total_reads) if read_records is None else (max(0, int(read_records[0])), min(total_reads, int(read_records[1])))
    reads = None
    if with_reads:
        nbytes = max(0, b - a) * RB
        reads = reads_out[:nbytes] if reads_out is not None else torch.empty(nbytes, dtype=torch.uint8, device=device)
    panels, segs, truth = [], [], []
    seg_off, seg_windows = [0], 0
    first = 0
    for c in range(spec.n_chrom):
        lo, hi = max(a, first), min(b, first + n_reads[c])
        need_reads = with_reads and lo < hi
        need_panel = with_panels and c in my
        need_segs = with_segments and (segment_chroms is None or c in segment_chroms)
        if not (need_reads or need_panel or need_segs):
            first += n_reads[c]
            seg_off.append(seg_off[-1])
            continue
        ch = make_chrom(spec, c, device)
        assert ch.n_reads == n_reads[c]
        if need_panel:
            panels.append(build_panel(spec, ch))
            truth.append(ch.truth.cpu().numpy())
        if need_segs:
            txt, nwin = segments_text(spec, ch)
            segs.append(txt)
            seg_windows += nwin
            del txt
        seg_off.append(seg_off[-1] + (int(segs[-1].numel()) if need_segs else 0))
        if need_reads:
            g = _gen(device, spec.seed, c, 5)
            h0 = haplotype_sequence(spec, ch, ch.truth[:, 0], g)
            h1 = haplotype_sequence(spec, ch, ch.truth[:, 1], g)
            haps = torch.cat([h0, h1])
            hap_off = (0, h0.numel())
            hap_len = (torch.tensor(h0.numel(), device=device), torch.tensor(h1.numel(), device=device))
            del h0, h1
            for chunk in range((lo - first) // READ_CHUNK, (hi - first + READ_CHUNK - 1) // READ_CHUNK):
                c_lo = first + chunk * READ_CHUNK
                n = min(READ_CHUNK, n_reads[c] - chunk * READ_CHUNK)
                txt = reads_chunk(spec, ch, haps, hap_off, hap_len, chunk, c_lo, n)
                s_lo, s_hi = max(lo, c_lo), min(hi, c_lo + n)
                reads[(s_lo - a) * RB:(s_hi - a) * RB] = txt[(s_lo - c_lo) * RB:(s_hi - c_lo) * RB]
                del txt
            del haps
        first += n_reads[c]
        del ch
    segments = torch.cat(segs) if with_segments and segs else None
    return Workload(spec, device, per, n_reads, my if with_panels else [], panels, segments, reads, (a, b), truth, seg_off, seg_windows)


def scaled(spec: Spec, n_variants: int, coverage: float | None = None) -> Spec:
    """The same per-column shape with fewer variants (tests)."""
    return replace(spec, n_variants=n_variants, coverage=spec.coverage if coverage is None else coverage)
