#!/bin/bash
# One batched GPU session (a gpurun call costs box time from the first second, so everything goes in one call).
# usage: scripts/gpu_round.sh <tag> [pytest|smoke|bench|bench_big|bench_hmm|launches|ncu|ncu_h64|sanitizer ...]
tag=$1; shift
out=gpurun_out; mkdir -p $out
for what in "$@"; do
case $what in
pytest)
  timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -150 > $out/pytest_gpu_$tag.log; tail -5 $out/pytest_gpu_$tag.log;;
smoke)
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; tail -3 $out/smoke_$tag.log;;
bench)
  timeout 900 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; tail -c 3000 $out/bench_$tag.json; tail -5 $out/bench_$tag.err;;
bench_big)
  for w in cfg3s h64s; do
    timeout 1200 python bench.py --steps 4 --warmup 3 --workload $w --no-cpu-baseline > $out/bench_${w}_$tag.json 2> $out/bench_${w}_$tag.err
    tail -c 1500 $out/bench_${w}_$tag.json; tail -3 $out/bench_${w}_$tag.err
  done;;
bench_hmm)
  timeout 1200 python scripts/bench_hmm.py > $out/bench_hmm_$tag.jsonl 2> $out/bench_hmm_$tag.err; cat $out/bench_hmm_$tag.jsonl; tail -3 $out/bench_hmm_$tag.err;;
launches)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/launches_$tag.csv \
     python bench.py --inflight 1 --steps 1 --warmup 1 --no-cpu-baseline > $out/launches_$tag.out 2>&1; tail -3 $out/launches_$tag.out;;
ncu)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"count_tile_kernel|basis_kernel|^scan_kernel|block_kernel|probe_parts" -s 10 -c 7 -o $out/prof_$tag \
     python bench.py --inflight 1 --steps 1 --warmup 1 --no-cpu-baseline > $out/ncu_$tag.out 2>&1; tail -3 $out/ncu_$tag.out;;
ncu_h64)
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"skeleton_kernel|block_kernel" -s 3 -c 2 -o $out/prof_h64_$tag \
     python bench.py --inflight 1 --steps 1 --warmup 1 --no-cpu-baseline --workload h64s > $out/ncu_h64_$tag.out 2>&1; tail -3 $out/ncu_h64_$tag.out;;
sanitizer)
  timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_hmm.py tests/test_gpu_counting.py -m gpu -q -p no:cacheprovider \
     -k "reference_vector or edge_cases or options or kmercounter_vectors or scan_and_sequential or partitioned" 2>&1 | tail -60 > $out/sanitizer_$tag.log; tail -8 $out/sanitizer_$tag.log;;
esac
done
