"""Host-side sharding logic of the multi-GPU path, exercised with world_size 2 over gloo on CPU.
The per-rank counting is done by the CPU oracle here (the GPU kernels are covered by the -m gpu tests); what is
tested is the partitioning (LPT chromosomes, record-aligned read shards) and that layout-identical tables whose
count arrays are all-reduced reproduce the single-process counts exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from synthdata import small as synth
from pangenie_b200.distributed import allreduce_counts, lpt_assign, record_ranges


class _HostCounter:
    """Stand-in for a device k-mer table in the host-logic test: counts interleaved with other data, moved through a
    contiguous exchange buffer by ranges exactly like pg_count_export_range / pg_count_import_range."""

    def __init__(self, counts: np.ndarray):
        self.table = np.zeros((len(counts), 4), np.int32)
        self.table[:, 2] = counts
        self.buf = None

    def capacity(self):
        return self.table.shape[0]

    def view(self, _counter, n):
        self.buf = torch.zeros(n, dtype=torch.int32)
        return self.buf

    def export_range(self, first, n):
        self.buf[:n] = torch.from_numpy(self.table[first:first + n, 2].copy())

    def import_range(self, first, n):
        self.table[first:first + n, 2] = self.buf[:n].numpy()


def test_lpt_assign_balances_and_is_deterministic():
    w = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51]
    for n in (1, 2, 4, 8):
        a = lpt_assign(w, n)
        assert sorted(sum(a, [])) == list(range(22))
        loads = [sum(w[i] for i in r) for r in a]
        assert max(loads) <= sum(w) / n + max(w) * 0.5
        assert a == lpt_assign(w, n)
    assert lpt_assign([5, 1], 4) == [[0], [1], [], []]


def test_record_ranges_fastq_and_fasta():
    wl = synth.make_workload(n_chrom=1, n_variants=60, n_haplotypes=4, coverage=3.0, seed=5)
    for text in (wl.reads_fastq, wl.segments_fasta):
        for n in (1, 2, 3, 8):
            rr = record_ranges(text, n)
            assert rr[0][0] == 0 and rr[-1][1] == len(text)
            for (a, b), (c, d) in zip(rr, rr[1:]):
                assert b == c
            for a, b in rr:
                if b > a:
                    assert text[a] in (ord("@"), ord(">")) and (a == 0 or text[a - 1] == 10)
    # quality lines starting with '@' must not be mistaken for record starts
    rec = b"@r\nACGT\n+\n@@@@\n"
    t = np.frombuffer(rec * 50, np.uint8)
    for a, b in record_ranges(t, 7):
        assert a % len(rec) == 0


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests import oracles
        import pangenie_b200 as pg
        lib = oracles.load_oracle()
        wl = synth.make_workload(n_chrom=3, n_variants=240, n_haplotypes=4, coverage=5.0, seed=9)
        probes = np.concatenate([p.kmer_codes for p in wl.panels] + [p.flank_codes for p in wl.panels])
        # every rank primes deterministically from the same segment file -> identical key sets; counts are exchanged
        # for a fixed probe list (stand-in for the layout-identical count arrays of the device tables)
        a, b = record_ranges(wl.reads_fastq, world)[rank]
        c = oracles.OracleCounter(lib, None, None, wl.k)
        c.feed(wl.segments_fasta, pg.PG_OP_PRIME)
        c.feed(wl.reads_fastq[a:b], pg.PG_OP_UPDATE)
        local = torch.from_numpy(c.lookup(probes).astype(np.int64))
        dist.all_reduce(local, op=dist.ReduceOp.SUM)
        full = oracles.OracleCounter(lib, wl.reads_fastq, wl.segments_fasta, wl.k)
        assert np.array_equal(local.numpy(), full.lookup(probes).astype(np.int64)), "sharded counts differ"
        # the exchange itself: the count arrays of layout-identical tables summed range by range through the exchange
        # buffer (capacity not a multiple of the chunk, several calls)
        rng = np.random.default_rng(100 + rank)
        mine_counts = rng.integers(0, 50, size=1000 + 4 * 7).astype(np.int32)
        hc = _HostCounter(mine_counts)
        calls = allreduce_counts(hc, world, chunk_slots=256, view=hc.view)
        want = sum(np.random.default_rng(100 + r).integers(0, 50, size=1000 + 4 * 7).astype(np.int32) for r in range(world))
        assert calls == 5 and np.array_equal(hc.table[:, 2], want) and not hc.table[:, [0, 1, 3]].any()
        # chromosome assignment: every chromosome genotyped exactly once across ranks
        mine = lpt_assign([p.n_variants for p in wl.panels], world)[rank]
        got = [None] * world
        dist.all_gather_object(got, mine)
        assert sorted(sum(got, [])) == [0, 1, 2]
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_sharded_counting_world_size_2_gloo(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
