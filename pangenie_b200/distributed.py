"""Single-sample sharding over the GPUs of one box.

Chromosomes are independent HMM chains and reads are independent units, so:
  * chromosomes are assigned to ranks by LPT (longest processing time first) on their variant counts,
    mirroring the reference's descending-size job order (reference src/graphbuilder.cpp:273-276);
  * the read file is cut into record-aligned byte ranges, one per rank;
  * every rank PRIMEs the k-mer table from the segment file itself and rewrites it into the canonical key layout
    (pg_count_canonicalize: a function of the key SET only), so all ranks hold the identical table without a broadcast;
    every rank UPDATEs its read shard; the count arrays are ALL-REDUCED - the one collective of the path.  Histogram peak,
    fill, emission and forward-backward run per rank on its own chromosomes with no further exchange.
The exchange buffer is exposed by pg_count_exchange_buffer and wrapped zero-copy as a torch tensor.
"""
from __future__ import annotations

import numpy as np


def lpt_assign(weights, n_ranks: int) -> list[list[int]]:
    """Longest-processing-time-first assignment of items (chromosomes) to ranks; deterministic."""
    order = sorted(range(len(weights)), key=lambda i: (-int(weights[i]), i))
    load = [0] * n_ranks
    out: list[list[int]] = [[] for _ in range(n_ranks)]
    for i in order:
        r = min(range(n_ranks), key=lambda q: (load[q], q))
        out[r].append(i)
        load[r] += int(weights[i])
    for lst in out:
        lst.sort()
    return out


def record_ranges(text: np.ndarray, n_shards: int) -> list[tuple[int, int]]:
    """Cuts a FASTA ('>') or 4-line FASTQ ('@') byte buffer into `n_shards` record-aligned ranges."""
    n = int(text.size)
    if n == 0:
        return [(0, 0)] * n_shards
    fastq = text[0] == ord("@")
    cuts = [0]
    for s in range(1, n_shards):
        p = n * s // n_shards
        # to the next line start
        while p < n and text[p - 1] != 10:
            p += 1
        while p < n:
            if fastq:
                # a record starts at a line beginning with '@' whose line+2 begins with '+'
                q, nl = p, 0
                while q < n and nl < 2:
                    if text[q] == 10:
                        nl += 1
                    q += 1
                if text[p] == ord("@") and q < n and text[q] == ord("+"):
                    break
            elif text[p] == ord(">"):
                break
            while p < n and text[p] != 10:
                p += 1
            p += 1
        cuts.append(min(p, n))
    cuts.append(n)
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return [(cuts[i], cuts[i + 1]) for i in range(n_shards)]


class _CudaArray:
    """Minimal __cuda_array_interface__ holder so torch.as_tensor can alias library-owned HBM."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


EXCHANGE_SLOTS = 1 << 28   # counts moved per all-reduce call (1 GiB of u32): large enough for full NVLink bus bandwidth


def exchange_tensor(counter, n_slots: int):
    """torch int32 view (no copy) of the counter's contiguous exchange buffer of `n_slots` counts."""
    import torch
    addr = counter.exchange_buffer(n_slots)
    return torch.as_tensor(_CudaArray(addr, n_slots, "<i4"), device="cuda")


def allreduce_counts(counter, world: int, group=None, chunk_slots: int = EXCHANGE_SLOTS, view=None, sync=None):
    """Adds the count arrays of all ranks position by position: THE collective of the sharded sample.

    The counts live interleaved with the keys in 64-byte buckets, so they pass through a contiguous exchange buffer in
    pieces of `chunk_slots` counts: export (device kernel) -> all-reduce (NCCL on torch's stream) -> import.  Every rank must
    hold the canonical key layout (KmerCounter.canonicalize after PRIME).  `view` / `sync` are injection points for the
    host-logic test (a numpy-backed counter and gloo)."""
    import torch.distributed as dist
    if world <= 1:
        return 0
    cap = counter.capacity()
    chunk = min(cap, max(4, chunk_slots & ~3))
    buf = (view or exchange_tensor)(counter, chunk)
    calls = 0
    for first in range(0, cap, chunk):
        n = min(chunk, cap - first)
        counter.export_range(first, n)          # synchronises the counter's stream: the buffer is complete
        dist.all_reduce(buf[:n], op=dist.ReduceOp.SUM, group=group)
        if sync is not None:
            sync()                               # NCCL runs on torch's stream: the sums must be complete before the import kernel
        counter.import_range(first, n)
        calls += 1
    return calls


def sharded_count(counter, reads, segments, rank: int, world: int, group=None, chunk_slots: int = EXCHANGE_SLOTS):
    """Every rank PRIMEs the full segment file into its own table and brings it into the canonical layout (identical on
    all ranks, no broadcast), UPDATEs with its own record-aligned shard of the reads, then the count arrays are summed
    over the ranks by all-reduce.  `counter` must have been created with the same max_distinct on every rank.
    Returns {"prime_ms", "update_ms", "allreduce_calls", "exchange_ms"}."""
    import time

    import torch
    from . import PG_OP_PRIME, PG_OP_UPDATE
    counter.feed(segments, PG_OP_PRIME)
    prime_ms = counter.last_ms()
    if world > 1:
        counter.canonicalize()
    update_ms = 0.0
    if reads is not None and len(reads):
        counter.feed(reads, PG_OP_UPDATE)
        update_ms = counter.last_ms()
    t0 = time.perf_counter()   # export / import synchronise the counter's stream and `sync` torch's: host time = device time here
    calls = allreduce_counts(counter, world, group, chunk_slots, sync=lambda: torch.cuda.current_stream().synchronize())
    return {"prime_ms": prime_ms, "update_ms": update_ms, "allreduce_calls": calls, "exchange_ms": 1e3 * (time.perf_counter() - t0)}
