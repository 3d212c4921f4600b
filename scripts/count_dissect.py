#!/usr/bin/env python
"""Timing dissection of the UPDATE pass (debug knob PG_COUNT_DEBUG of count_tile_kernel<.,SCATTER>): the full configs[2] table,
the first 12 M reads.  Prints scatter / probe ms for each knob value."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1:
    import torch
    import pangenie_b200 as pg
    from synthdata import large
    spec = large.CONFIGS["cfg3"]
    wl = large.make_workload(spec, "cuda", with_panels=False, read_records=(0, 12_000_000))
    c = pg.KmerCounter(None, None, spec.k, max_distinct=int(wl.segments.numel()))
    c.feed(wl.segments, pg.PG_OP_PRIME)
    for rep in range(3):
        c.feed(wl.reads, pg.PG_OP_UPDATE)
        ms, (pms, n) = c.last_ms(), c.last_probe_ms()
    print(f"debug {sys.argv[1]:>3s}: update {ms:7.1f} ms  probe {pms:7.1f} ms ({n} passes)  scatter {ms - pms:7.1f} ms  for {wl.reads.numel() / 1e9:.2f} GB")
else:
    for dbg in ("0", "16", "32", "64", "96", "112"):
        subprocess.run([sys.executable, __file__, dbg], env=dict(os.environ, PG_COUNT_DEBUG=dbg))
