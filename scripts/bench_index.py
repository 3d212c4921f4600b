#!/usr/bin/env python
"""Index stage at scale: unique-k-mer selection for V bubbles of the configs[2] mix (90 % SNPs, 8 % indels of
1-50 bp, 2 % tri-allelic) on the device (pg_unique_kmers_compute) next to the CPU restatement of
StepwiseUniqueKmerComputer::compute_unique_kmers (one thread, as the reference runs it per chromosome) on a sample.

  python scripts/bench_index.py [--variants 1000000] [--paths 33] [--cpu-sample 20000]
Prints one JSON line: device kernel time (CUDA events inside the library), whole call (upload + kernel + download + host
compaction), variants/s, the CPU restatement's variants/s on the sample and whether the two agree on it."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_inputs(V, P, k, seed):
    """Flat pg_variants arrays of one synthetic chromosome + its path-segment FASTA (numpy, vectorised)."""
    rng = np.random.default_rng(seed)
    spacing = rng.integers(100, 1100, V)
    pos = np.cumsum(spacing) + 2 * k
    G = int(pos[-1]) + 1200
    ref = rng.integers(0, 4, G, dtype=np.uint8)
    kind = rng.random(V)
    reflen = np.where(kind < 0.90, 1, np.where(kind < 0.94, rng.integers(2, 51, V), 1)).astype(np.int64)   # 4 % deletions
    n_alt = np.where(kind >= 0.98, 2, 1)
    altlen = np.where((kind >= 0.94) & (kind < 0.98), rng.integers(2, 51, V), 1).astype(np.int64)        # 4 % insertions
    letters = np.frombuffer(b"ACGT", np.uint8)
    seqs, soff, aoff, undef = [], [0], [0], []
    lefts, rights, loff, roff = [], [], [0], [0]
    seg = []
    prev_end = 0
    ends = pos + reflen
    for v in range(V):
        s, e = int(pos[v]), int(ends[v])
        left, right = ref[s - (k - 1):s], ref[e:e + k - 1]
        alleles = [ref[s:e]]
        for a in range(int(n_alt[v])):
            alt = rng.integers(0, 4, int(altlen[v]), dtype=np.uint8)
            if len(alt) == len(alleles[0]) == 1 and alt[0] == alleles[0][0]:
                alt = (alt + 1 + a) % 4
            alleles.append(alt.astype(np.uint8))
        seg.append(b">c_reference_%d\n" % s + letters[ref[prev_end:s]].tobytes() + b"\n")
        for a, al in enumerate(alleles):
            full = np.concatenate([left, al, right])
            seqs.append(full)
            soff.append(soff[-1] + len(full))
            undef.append(0)
            seg.append(b">c_%d_%d\n" % (s, a) + letters[full].tobytes() + b"\n")
        aoff.append(aoff[-1] + len(alleles))
        lo = max(s - 2 * k, prev_end)
        nxt = int(pos[v + 1]) if v + 1 < V else G
        lefts.append(ref[lo:s]); loff.append(loff[-1] + s - lo)
        hi = min(e + 2 * k, nxt)
        rights.append(ref[e:hi]); roff.append(roff[-1] + hi - e)
        prev_end = e
    seg.append(b">c_reference_end\n" + letters[ref[prev_end:]].tobytes() + b"\n")
    na = np.diff(np.array(aoff))
    p2a = (rng.integers(0, 1 << 30, (V, P)) % na[:, None]).astype(np.uint16)
    p2a[:, 0] = 0
    flat = dict(k=k, n_variants=V, n_paths=P, positions=pos.astype(np.uint64), end_positions=ends.astype(np.uint64),
                path_to_allele=p2a.reshape(-1), allele_offsets=np.array(aoff, np.uint32), allele_undefined=np.array(undef, np.uint8),
                seq_offsets=np.array(soff, np.uint64), seq=letters[np.concatenate(seqs)],
                left_offsets=np.array(loff, np.uint64), left_seq=letters[np.concatenate(lefts)],
                right_offsets=np.array(roff, np.uint64), right_seq=letters[np.concatenate(rights)])
    return flat, b"".join(seg)


def head(flat, n):
    """The first n variants of `flat` (same graph counts: a sample of the same chromosome)."""
    a1 = int(flat["allele_offsets"][n])
    out = dict(flat)
    out.update(n_variants=n, positions=flat["positions"][:n], end_positions=flat["end_positions"][:n],
               path_to_allele=flat["path_to_allele"][:n * flat["n_paths"]], allele_offsets=flat["allele_offsets"][:n + 1],
               allele_undefined=flat["allele_undefined"][:a1], seq_offsets=flat["seq_offsets"][:a1 + 1],
               left_offsets=flat["left_offsets"][:n + 1], right_offsets=flat["right_offsets"][:n + 1])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", type=int, default=1_000_000)
    ap.add_argument("--paths", type=int, default=33)
    ap.add_argument("--cpu-sample", type=int, default=20_000)
    args = ap.parse_args()
    import pangenie_b200 as pg
    from tests import oracles
    k = 31
    t0 = time.time()
    flat, seg = make_inputs(args.variants, args.paths, k, 20260925)
    t_gen = time.time() - t0
    counts = pg.KmerCounter(kmer_size=k, max_distinct=int(len(seg) * 1.05))
    t0 = time.time()
    counts.feed(seg, pg.PG_OP_COUNT)
    t_count = time.time() - t0
    best = None
    for _ in range(3):
        t0 = time.time()
        sel = pg.UniqueKmerSelection(counts, flat)
        wall = time.time() - t0
        ms, n = sel.stats()
        if best is None or ms < best[0]:
            best = (ms, wall, n)
        pan = sel.panel()
        sel.close()
    n = args.cpu_sample
    lib = oracles.load_oracle()
    oc = oracles.OracleCounter(lib, k=k)
    oc.feed(seg, pg.PG_OP_COUNT, threads=os.cpu_count() or 1)
    sub = head(flat, n)
    t0 = time.time()
    want = oracles.oracle_unique_kmers(lib, oc, sub)
    t_cpu = time.time() - t0
    K = int(want.kmer_offsets[-1])
    same = bool(np.array_equal(want.kmer_codes, pan.kmer_codes[:K]) and np.array_equal(want.kmer_offsets, pan.kmer_offsets[:n + 1]) and
                np.array_equal(want.flank_codes, pan.flank_codes[:int(want.flank_offsets[-1])]) and
                np.array_equal(want.allele_kmer_mask, pan.allele_kmer_mask[:len(want.allele_kmer_mask)]))
    print(json.dumps({"stage": "index: unique-k-mer selection", "variants": args.variants, "paths": args.paths,
                      "kmers_enumerated": best[2], "selected_kmers": int(pan.kmer_offsets[-1]), "flank_kmers": int(pan.flank_offsets[-1]),
                      "kernel_ms": best[0], "call_s": best[1], "variants_per_s_kernel": args.variants / (best[0] * 1e-3),
                      "variants_per_s_call": args.variants / best[1], "graph_count_s": t_count, "segment_bytes": len(seg),
                      "cpu_restatement": {"variants": n, "seconds": t_cpu, "variants_per_s": n / t_cpu, "threads": 1,
                                          "identical_to_device": same}, "input_generation_s": t_gen}))


if __name__ == "__main__":
    main()
