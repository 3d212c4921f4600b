"""Test infrastructure: the INPUT side of the reference's index stage, i.e. what
`StepwiseUniqueKmerComputer` receives — per variant bubble the allele sequences with their flanks, the paths, and the
reference sequence left / right of the bubble.

Two sources:
* `read_graph_cereal` decodes a `<prefix>_<chrom>_Graph.cereal` written by the real PanGenie-index (cereal
  BinaryOutputArchive: raw little-endian, size_t -> u64, string / vector = u64 length + elements, non-polymorphic
  shared_ptr = u32 pointer id with the MSB set when the object follows).  Field order: reference src/graph.hpp:81-84
  (`fasta_reader, chromosome, kmer_size, add_reference, variants_deleted, variants, variant_ids`),
  src/fastareader.hpp:41-43, src/dnasequence.hpp:49-51 (two bases per byte, high nibble first, 4 = undefined),
  src/variant.hpp:178-180.
* `graph_from_vcf` builds the same records from a VCF + reference FASTA for variants that are at least k apart
  (reference src/graphbuilder.cpp:70-288 without the merging of close variants into bubbles; enough for demo/).

`flatten` turns the records into the flat arrays `pg_unique_kmers_compute` takes (include/pangenie_b200.h).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

_DEC = "ACGTN"


class _R:
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

    def dna(self):  # DnaSequence::serialize: vector<unsigned char> sequence, bool even_length, bool is_undefined
        n = self.take("Q")
        raw = self.d[self.o:self.o + n]
        self.o += n
        even, undefined = self.take("B"), self.take("B")
        out = []
        for b in raw:
            out.append(_DEC[min(b >> 4, 4)])
            out.append(_DEC[min(b & 15, 4)])
        if not even:
            out.pop()
        s = "".join(out)
        assert ("N" in s) == bool(undefined)
        return s


@dataclass
class Bubble:
    """One Variant object of the reference (src/variant.hpp:172-196) reduced to what the index stage reads."""
    chromosome: str
    start: int
    end: int                          # Variant::get_end_position (src/variant.cpp:207-216)
    alleles: list                     # get_allele_sequence(a) WITH flanks (src/variant.cpp:181-201)
    undefined: list                   # is_undefined_allele(a) (src/variant.cpp:625-632: flanks do not count)
    paths: list                       # get_allele_on_path(p)
    left_overhang: str = ""           # Graph::get_left_overhang(v, 2k) (src/graph.cpp:554-572)
    right_overhang: str = ""          # Graph::get_right_overhang(v, 2k) (src/graph.cpp:574-592)


@dataclass
class RefGraph:
    chromosome: str
    kmer_size: int
    add_reference: bool
    reference: str
    bubbles: list = field(default_factory=list)

    def set_overhangs(self):
        k2 = 2 * self.kmer_size
        n = len(self.bubbles)
        for i, b in enumerate(self.bubbles):
            prev_end = self.bubbles[i - 1].end if i > 0 else 0
            lo = max(b.start - k2, prev_end) if b.start >= k2 else prev_end   # size_t arithmetic: start < 2k wraps, then clamps
            b.left_overhang = self.reference[lo:b.start]
            nxt = self.bubbles[i + 1].start if i + 1 < n else len(self.reference)
            b.right_overhang = self.reference[b.end:min(b.end + k2, nxt)]

    def segments_fasta(self) -> str:
        """GraphBuilder::write_path_segments (src/graphbuilder.cpp:293-353) for this chromosome."""
        out, prev = [], 0
        for b in self.bubbles:
            out.append(f">{self.chromosome}_reference_{b.start}\n{self.reference[prev:b.start]}\n")
            for a, s in enumerate(b.alleles):
                out.append(f">{self.chromosome}_{b.start}_{a}\n{s}\n")
            prev = b.end
        out.append(f">{self.chromosome}_reference_end\n{self.reference[prev:]}\n")
        return "".join(out)


def read_graph_cereal(path: str) -> RefGraph:
    r = _R(open(path, "rb").read())
    seqs = {}
    for _ in range(r.take("Q")):          # FastaReader::name_to_sequence : map<string, shared_ptr<DnaSequence>>
        name = r.string()
        pid = r.take("I")
        assert pid & 0x80000000
        seqs[name] = r.dna()
    chrom = r.string()
    k = r.take("Q")
    add_ref, deleted = r.take("B"), r.take("B")
    assert not deleted
    g = RefGraph(chrom, k, bool(add_ref), seqs[chrom])
    for _ in range(r.take("Q")):          # vector<shared_ptr<Variant>>
        pid = r.take("I")
        assert pid & 0x80000000
        left, right = r.dna(), r.dna()
        inner = [r.dna() for _ in range(r.take("Q"))]
        vchrom = r.string()
        start = r.take("Q")
        allele_seqs = [[r.dna() for _ in range(r.take("Q"))] for _ in range(r.take("Q"))]
        combos = [[r.take("H") for _ in range(r.take("Q"))] for _ in range(r.take("Q"))]
        _uncovered = [[r.take("H") for _ in range(r.take("Q"))] for _ in range(r.take("Q"))]
        paths = [r.take("H") for _ in range(r.take("Q"))]
        flanks_added = r.take("B")
        assert flanks_added, "PanGenie-index serialises the graph with flanks added"
        end = start + sum(len(a[0]) for a in allele_seqs) + sum(len(x) for x in inner[:len(allele_seqs) - 1])
        alleles, undefined = [], []
        for combo in combos:
            s = left
            for i, ai in enumerate(combo):
                s += allele_seqs[i][ai]
                if i < len(combo) - 1:
                    s += inner[i]
            alleles.append(s + right)
            undefined.append(any("N" in allele_seqs[i][ai] for i, ai in enumerate(combo)))
        g.bubbles.append(Bubble(vchrom, start, end, alleles, undefined, paths))
    for _ in range(r.take("Q")):          # variant_ids: vector<vector<string>>
        for _i in range(r.take("Q")):
            r.string()
    assert r.o == len(r.d), "trailing bytes in Graph.cereal"
    g.set_overhangs()
    return g


def read_fasta(path: str) -> dict:
    """FastaReader::parse_file (src/fastareader.cpp): name = first word of the header, sequence upper-cased by DnaSequence
    (non-ACGT -> N)."""
    seqs, name = {}, None
    for line in open(path):
        line = line.rstrip("\r\n")
        if line.startswith(">"):
            name = line[1:].split()[0]
            seqs[name] = []
        elif name is not None:
            seqs[name].append("".join(c if c in "ACGT" else "N" for c in line.upper()))
    return {n: "".join(s) for n, s in seqs.items()}


def graph_from_vcf(vcf_path: str, fasta_path: str, k: int = 31, add_reference: bool = True) -> dict:
    """-> {chromosome: RefGraph}; variants closer than k to their predecessor are not supported here (the reference merges
    them into one bubble, src/graphbuilder.cpp:186-206)."""
    ref = read_fasta(fasta_path)
    graphs = {}
    prev_end = {}
    for line in open(vcf_path):
        if line.startswith("#") or not line.strip():
            continue
        t = line.rstrip("\n").split("\t")
        chrom, pos, refa, alts = t[0], int(t[1]) - 1, t[3].upper(), t[4].upper().split(",")
        g = graphs.setdefault(chrom, RefGraph(chrom, k, add_reference, ref[chrom]))
        assert pos - prev_end.get(chrom, -10 ** 9) >= k, "close variants would be merged by the reference"
        paths = [0] if add_reference else []
        for gt in t[9:]:
            for a in gt.replace("/", "|").split("|"):
                assert a != ".", "undefined genotypes are not supported by this helper"
                paths.append(int(a))
        end = pos + len(refa)
        assert g.reference[pos:end] == refa
        left, right = g.reference[pos - (k - 1):pos], g.reference[end:end + k - 1]
        alleles = [left + a + right for a in [refa] + alts]
        g.bubbles.append(Bubble(chrom, pos, end, alleles, ["N" in a for a in [refa] + alts], paths))
        prev_end[chrom] = end
    for g in graphs.values():
        g.set_overhangs()
    return graphs


def flatten(g: RefGraph) -> dict:
    """Flat arrays of one chromosome for pg_unique_kmers_compute / the oracle."""
    V = len(g.bubbles)
    P = len(g.bubbles[0].paths) if V else 0
    aoff, soff, loff, roff = [0], [0], [0], [0]
    seq, lseq, rseq, undef = [], [], [], []
    for b in g.bubbles:
        for s, u in zip(b.alleles, b.undefined):
            seq.append(s)
            soff.append(soff[-1] + len(s))
            undef.append(1 if u else 0)
        aoff.append(aoff[-1] + len(b.alleles))
        lseq.append(b.left_overhang)
        loff.append(loff[-1] + len(b.left_overhang))
        rseq.append(b.right_overhang)
        roff.append(roff[-1] + len(b.right_overhang))
    return dict(
        k=g.kmer_size, n_variants=V, n_paths=P,
        positions=np.array([b.start for b in g.bubbles], np.uint64),
        path_to_allele=np.array([a for b in g.bubbles for a in b.paths], np.uint16),
        allele_offsets=np.array(aoff, np.uint32),
        allele_undefined=np.array(undef, np.uint8),
        seq_offsets=np.array(soff, np.uint64),
        seq=np.frombuffer("".join(seq).encode(), np.uint8).copy(),
        left_offsets=np.array(loff, np.uint64),
        left_seq=np.frombuffer("".join(lseq).encode(), np.uint8).copy(),
        right_offsets=np.array(roff, np.uint64),
        right_seq=np.frombuffer("".join(rseq).encode(), np.uint8).copy(),
    )
