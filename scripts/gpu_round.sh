#!/bin/bash
# One batched GPU session (a gpurun call costs box time from the first second, so everything goes in one call).
# usage: scripts/gpu_round.sh <tag> [pytest|smoke|bench|bench_ref|bench_cfg2|bench_hmm|launches|ncu|ncu_hmm|sanitizer|racecheck ...]
tag=$1; shift
out=gpurun_out; mkdir -p $out
for what in "$@"; do
case $what in
pytest)
  timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | tail -150 > $out/pytest_gpu_$tag.log; tail -5 $out/pytest_gpu_$tag.log;;
pytest_new)
  timeout 1500 python -m pytest tests/test_gpu_at_size.py -m gpu -q -x --timeout 900 -p no:cacheprovider 2>&1 | tail -150 > $out/pytest_new_$tag.log; tail -15 $out/pytest_new_$tag.log;;
smoke)
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; tail -3 $out/smoke_$tag.log;;
bench)
  timeout 1500 python bench.py --steps ${STEPS:-5} --warmup 3 > $out/bench_$tag.json 2> $out/bench_$tag.err; tail -c 6000 $out/bench_$tag.json; tail -5 $out/bench_$tag.err;;
bench_ref)
  timeout 900 python bench.py --impl reference --steps 2 --warmup 0 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err; tail -c 3000 $out/bench_ref_$tag.json; tail -5 $out/bench_ref_$tag.err;;
bench_cfg2)
  timeout 900 python bench.py --workload cfg2 --steps 20 --warmup 3 > $out/bench_cfg2_$tag.json 2> $out/bench_cfg2_$tag.err; tail -c 3000 $out/bench_cfg2_$tag.json; tail -5 $out/bench_cfg2_$tag.err;;
bench_hmm)
  timeout 1200 python scripts/bench_hmm.py > $out/bench_hmm_$tag.jsonl 2> $out/bench_hmm_$tag.err; cat $out/bench_hmm_$tag.jsonl; tail -3 $out/bench_hmm_$tag.err;;
launches)
  # every launch of one step of the default workload with its device time (cold-cache, serialised: shares, not absolutes)
  # (ncu writes its log at exit: bound the number of profiled launches and the wall time, or a budget-clamped call loses everything)
  PG_BENCH_MIN_WARMUP=1 PG_BENCH_E2E_STEPS=0 timeout ${LAUNCH_TIMEOUT:-600} ncu --metrics gpu__time_duration.sum --clock-control none -c ${NLAUNCH:-1500} --csv --log-file $out/launches_$tag.csv \
     python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --workload ${WL:-cfg3s} > $out/launches_$tag.out 2>&1; tail -3 $out/launches_$tag.out;;
ncu)
  PG_BENCH_MIN_WARMUP=1 PG_BENCH_E2E_STEPS=0 timeout 1500 ncu --set full --clock-control none --import-source on \
     -k regex:"count_tile_kernel|probe_parts|block_kernel|skeleton_kernel" -s ${SKIP:-40} -c ${CNT:-8} -o $out/prof_$tag \
     python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --workload ${WL:-cfg3s} > $out/ncu_$tag.out 2>&1; tail -3 $out/ncu_$tag.out;;
ncu_probe)
  PG_BENCH_MIN_WARMUP=1 PG_BENCH_E2E_STEPS=0 timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"probe_parts" -s 1 -c 2 -o $out/prof_probe_$tag \
     python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --workload ${WL:-cfg3} > $out/ncu_probe_$tag.out 2>&1; tail -3 $out/ncu_probe_$tag.out;;
ncu_scatter)
  PG_BENCH_MIN_WARMUP=1 PG_BENCH_E2E_STEPS=0 timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"count_tile_kernel" -s 30 -c 2 -o $out/prof_scatter_$tag \
     python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --workload ${WL:-cfg3} > $out/ncu_scatter_$tag.out 2>&1; tail -3 $out/ncu_scatter_$tag.out;;
ncu_skel)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"skeleton" -c 2 -o $out/prof_skel_$tag \
     python scripts/bench_hmm.py --haplotypes ${H:-64} --variants 100000 --repeat 1 > $out/ncu_skel_$tag.out 2>&1; tail -3 $out/ncu_skel_$tag.out;;
sweep_slices)
  for kb in 16384 24576 32768 49152 65536 98304; do
    PG_COUNT_PART_KB=$kb PG_BENCH_E2E_STEPS=0 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > $out/bench_slice${kb}_$tag.json 2> $out/bench_slice${kb}_$tag.err
    python - <<EOF
import json
d=json.loads([l for l in open("$out/bench_slice${kb}_$tag.json") if l.startswith("{")][0])
u=d["roofline"]["update_pass"]
print("slice_kb", $kb, "step_ms", round(d["ms_per_step"],1), "update_ms", round(u["ms"],1), "probe_ms", round(u["probe_ms"],1), "passes", u["probe_passes"])
EOF
  done;;
sweep_items)
  for cfg in ${SWEEP:-"16384 0" "24576 0" "40960 0" "65536 0" "98304 0"}; do
    set -- $cfg
    PG_COUNT_PART_KB=$1 PG_COUNT_PROBE_ITEM=$2 PG_BENCH_E2E_STEPS=0 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > $out/bench_sw_$1_$2_$tag.json 2> $out/bench_sw_$1_$2_$tag.err
    python - <<EOF
import json
d=json.loads([l for l in open("$out/bench_sw_$1_$2_$tag.json") if l.startswith("{")][0])
u=d["roofline"]["update_pass"]
print("slice_kb", $1, "item", $2, "step_ms", round(d["ms_per_step"],1), "update_ms", round(u["ms"],1), "probe_ms", round(u["probe_ms"],1), "scatter_ms", round(u["ms"]-u["probe_ms"],1), "passes", u["probe_passes"])
EOF
  done;;
direct)
  PG_COUNT_PART_KB=0 PG_BENCH_E2E_STEPS=0 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > $out/bench_direct_$tag.json 2> $out/bench_direct_$tag.err
  python - <<EOF
import json
d=json.loads([l for l in open("$out/bench_direct_$tag.json") if l.startswith("{")][0])
u=d["roofline"]["update_pass"]
print("direct (no partitioning): step_ms", round(d["ms_per_step"],1), "update_ms", round(u["ms"],1))
EOF
  ;;
hmm_cluster)
  for c in 0 2 4; do
    PG_SKELETON_CLUSTER=$c timeout 900 python scripts/bench_hmm.py --haplotypes 32 64 --variants 400000 --repeat 2 > $out/bench_hmm_cluster${c}_$tag.jsonl 2> $out/bench_hmm_cluster${c}_$tag.err
    python - <<EOF
import json
for l in open("$out/bench_hmm_cluster${c}_$tag.jsonl"):
    d=json.loads(l); print("cluster", $c, "H", d["haplotypes"], "skeleton_ms", round(d["skeleton_ms"],2), "blocks_ms", round(d["blocks_ms"],2))
EOF
    tail -2 $out/bench_hmm_cluster${c}_$tag.err
  done;;
ncu_hmm)
  PG_BENCH_MIN_WARMUP=1 PG_BENCH_E2E_STEPS=0 timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"skeleton_kernel|block_kernel" -s 1 -c 2 -o $out/prof_hmm_$tag \
     python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --workload ${WL:-cfg3s} > $out/ncu_hmm_$tag.out 2>&1; tail -3 $out/ncu_hmm_$tag.out;;
sanitizer)
  timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_hmm.py tests/test_gpu_counting.py tests/test_gpu_at_size.py tests/test_index_build.py tests/test_demo_config0.py -m gpu -q -p no:cacheprovider \
     -k "reference_vector or edge_cases or options or kmercounter_vectors or scan_and_sequential or partitioned or canonical or long_headers or subsets or lean or device_selection or demo" 2>&1 | tail -60 > $out/sanitizer_$tag.log; tail -8 $out/sanitizer_$tag.log;;
racecheck)
  timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_hmm.py tests/test_gpu_counting.py tests/test_index_build.py -m gpu -q -p no:cacheprovider \
     -k "reference_vector or kmercounter_vectors or scan_and_sequential or partitioned or lean or reference_index_fixture" 2>&1 | tail -60 > $out/racecheck_$tag.log; tail -8 $out/racecheck_$tag.log;;
bench_index)
  timeout 900 python scripts/bench_index.py --variants ${IDXV:-1000000} > $out/bench_index_$tag.json 2> $out/bench_index_$tag.err; cat $out/bench_index_$tag.json; tail -3 $out/bench_index_$tag.err;;
esac
done
