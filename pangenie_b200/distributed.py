"""Single-sample sharding over the GPUs of one box.

Chromosomes are independent HMM chains and reads are independent units, so:
  * chromosomes are assigned to ranks by LPT (longest processing time first) on their variant counts,
    mirroring the reference's descending-size job order (reference src/graphbuilder.cpp:273-276);
  * the read file is cut into record-aligned byte ranges, one per rank;
  * rank 0 PRIMEs the k-mer table from the segment file and BROADCASTS the key array (one NCCL call), so every
    rank holds a layout-identical table; every rank UPDATEs its read shard; the count arrays are ALL-REDUCED
    (one NCCL call).  No other collective: histogram peak, fill, emission and forward-backward run per rank on
    its own chromosomes.
The device arrays are exposed by pg_count_device_arrays and wrapped zero-copy as torch tensors.
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


def counter_tensors(counter):
    """(slots int64[2*cap], counts int32[cap]) torch views of a KmerCounter's device arrays (no copy).
    `slots` is the table itself (16-byte slots); `counts` is the contiguous staging array of the all-reduce."""
    import torch
    sp, cp, cap = counter.device_arrays()
    slots = torch.as_tensor(_CudaArray(sp, 2 * cap, "<i8"), device="cuda")
    counts = torch.as_tensor(_CudaArray(cp, cap, "<i4"), device="cuda")
    return slots, counts


def sharded_count(counter, reads, segments, rank: int, world: int, group=None):
    """PRIME on rank 0 + broadcast keys, UPDATE the local read shard, all-reduce counts.

    `reads` is this rank's record-aligned shard (host numpy / pinned torch / cuda torch uint8); `segments` is
    only read on rank 0.  `counter` must have been created with the same max_distinct on every rank.
    """
    import torch
    import torch.distributed as dist
    from . import PG_OP_PRIME, PG_OP_UPDATE
    slots, counts = counter_tensors(counter)
    if rank == 0:
        counter.feed(segments, PG_OP_PRIME)
    on_gpu = slots.is_cuda
    if world > 1:
        dist.broadcast(slots, src=0, group=group)
        if on_gpu:
            # NCCL runs on torch's stream, the counter on its own non-blocking stream: the keys must have
            # arrived before the UPDATE kernels probe them
            torch.cuda.current_stream().synchronize()
    if reads is not None and len(reads):
        counter.feed(reads, PG_OP_UPDATE)
    if world > 1:
        counter.export_counts()
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
        if on_gpu:
            torch.cuda.current_stream().synchronize()  # the reduced counts must be complete before the import kernel
        counter.import_counts()
    return counter
