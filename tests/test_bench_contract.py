"""bench.py's reference arm (CPU only) prints ONE JSON line with the keys the driver reads, loads nothing of the product
library, and extrapolates a bounded sub-sample the way its `sample` string says."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra, env=None):
    e = dict(os.environ, PG_BENCH_REF_BUDGET_S="1", PG_BENCH_REF_GEN_CPU="1", **(env or {}))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *extra],
                         capture_output=True, text=True, env=e, cwd=ROOT, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    d = _run("--workload", "tiny", "--steps", "2", "--warmup", "0")
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "variants/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and "600 variants" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"] > 0
    assert "nothing extrapolated" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "variants/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["steps"] == 2 and 1 <= d["steps_executed"] <= 2


def test_default_workloads_follow_baseline_json():
    sys.path.insert(0, ROOT)
    import bench
    from synthdata import large
    assert bench.default_workload(1) == "cfg3" and bench.default_workload(2) == "cfg3"
    assert bench.default_workload(4) == "cfg4" and bench.default_workload(8) == "cfg4"
    for name, idx in (("cfg2", 1), ("cfg3", 2), ("cfg4", 3), ("cfg5", 4)):
        assert f"configs[{idx}]" in large.CONFIGS[name].text
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    assert "1M variants" in base[2] and large.CONFIGS["cfg3"].n_variants == 1_000_000 and large.CONFIGS["cfg3"].n_haplotypes == 32
    assert "5M variants" in base[3] and large.CONFIGS["cfg4"].n_variants == 5_000_000 and large.CONFIGS["cfg4"].n_haplotypes == 64
    # identical config dicts in both arms whatever the number of ranks
    assert bench.make_config(large.CONFIGS["cfg3"], 1) == bench.make_config(large.CONFIGS["cfg3"], 8)


def test_makespan_and_sample_choice():
    sys.path.insert(0, ROOT)
    import bench
    from synthdata import large
    assert bench.makespan([3, 3, 3, 3], 2) == 6 and bench.makespan([5, 1, 1, 1], 2) == 5 and bench.makespan([2, 2], 8) == 2
    spec = large.CONFIGS["cfg3"]
    reads = [int(30 * 600 * v / 150) for v in large.variants_per_chrom(spec)]
    chroms = bench.choose_sample_chroms(spec, reads, large.record_bytes(spec))
    assert chroms == [20, 21]   # the two smallest autosomes (chr21, chr22): 1.2 GB of the 38 GB


def test_reference_arm_extrapolates_a_sub_sample():
    # 22 chromosomes with a budget that only admits the smallest ones: the line must say what was extrapolated
    sys.path.insert(0, ROOT)
    import bench
    import torch
    from synthdata import large
    spec = large.scaled(large.CONFIGS["cfg3"], 4400, coverage=4.0)
    orig = bench.choose_sample_chroms
    bench.choose_sample_chroms = lambda s, r, rb, budget_bytes=0: orig(s, r, rb, 3e5)
    try:
        sample = bench.generate_cpu_sample(spec, torch.device("cpu"))
    finally:
        bench.choose_sample_chroms = orig
    assert 1 <= len(sample["chroms"]) < 22 and sample["read_bytes_total"] > len(sample["reads"]) > 0
    r = bench.cpu_reference(spec, sample, 2)
    assert "extrapolated: PRIME x" in r["sample"] and r["value"] > 0
    tot = r["seconds_whole_sample"]
    assert np.isclose(tot["total"], tot["prime"] + tot["update"] + tot["histogram"] + tot["fill"] + tot["hmm"])
    assert tot["update"] > r["seconds_measured"]["update"]
