"""Two-GPU sharded run (NCCL) against the single-GPU result; skipped with fewer than two devices."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import pangenie_b200 as pg
        from synthdata import small as synth
        from pangenie_b200.distributed import lpt_assign, record_ranges, sharded_count
        wl = synth.make_workload(n_chrom=4, n_variants=1600, n_haplotypes=8, coverage=8.0, seed=31)
        mine = lpt_assign([p.n_variants for p in wl.panels], world)[rank]
        panels = [wl.panels[i] for i in mine]
        a, b = record_ranges(wl.reads_fastq, world)[rank]
        eng = pg.Engine(rank)
        counter = pg.KmerCounter(None, None, wl.k, max_distinct=len(wl.segments_fasta), device=rank)
        sharded_count(counter, wl.reads_fastq[a:b].copy(), wl.segments_fasta, rank, world)
        res = eng.load(panels)
        peak = eng.run_counted(counter, True, 0.01, recombrate=1.26, effective_N=1e-5)
        eng.fetch()
        np.savez(os.path.join(tmp, f"r{rank}.npz"), peak=peak, mine=np.array(mine),
                 **{f"lik{i}": r.likelihoods for i, r in zip(mine, res)}, **{f"gt{i}": r.genotype for i, r in zip(mine, res)},
                 **{f"cnt{i}": p.kmer_counts for i, p in zip(mine, panels)})
    finally:
        dist.destroy_process_group()


def test_two_gpu_sharded_run_matches_single_gpu(tmp_path, engine):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from synthdata import small as synth
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    wl = synth.make_workload(n_chrom=4, n_variants=1600, n_haplotypes=8, coverage=8.0, seed=31)
    want, peak = engine.genotype_run(wl.reads_fastq, wl.segments_fasta, wl.panels, k=wl.k, recombrate=1.26, effective_N=1e-5)
    seen = []
    for r in range(2):
        z = np.load(tmp_path / f"r{r}.npz")
        assert int(z["peak"]) == peak
        for i in z["mine"]:
            seen.append(int(i))
            assert np.array_equal(z[f"cnt{i}"], wl.panels[i].kmer_counts)          # integer work: bit-exact
            assert np.array_equal(z[f"gt{i}"], want[i].genotype)
            assert np.array_equal(z[f"lik{i}"], want[i].likelihoods)               # same kernels, same inputs
    assert sorted(seen) == [0, 1, 2, 3]
