#!/bin/bash
# One batched GPU session (gpurun charges ~10 min per call, so everything goes in one call).
# usage: scripts/gpu_round.sh <tag> [pytest|smoke|bench|launches|sanitizer|ncu ...]
tag=$1; shift
out=gpurun_out; mkdir -p $out
for what in "$@"; do
case $what in
pytest)
  timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -150 > $out/pytest_gpu_$tag.log; tail -5 $out/pytest_gpu_$tag.log;;
smoke)
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; tail -3 $out/smoke_$tag.log;;
bench)
  timeout 900 python bench.py --steps 5 --warmup 3 > $out/bench_$tag.json 2> $out/bench_$tag.err; tail -c 3000 $out/bench_$tag.json; tail -5 $out/bench_$tag.err;;
bench_big)
  timeout 1200 python bench.py --steps 3 --warmup 3 --workload cfg3s > $out/bench_cfg3s_$tag.json 2> $out/bench_cfg3s_$tag.err; tail -c 3000 $out/bench_cfg3s_$tag.json; tail -5 $out/bench_cfg3s_$tag.err
  timeout 1200 python bench.py --steps 3 --warmup 3 --workload h64s > $out/bench_h64s_$tag.json 2> $out/bench_h64s_$tag.err; tail -c 3000 $out/bench_h64s_$tag.json; tail -5 $out/bench_h64s_$tag.err;;
launches)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv \
     python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $out/launches_$tag.out 2>&1; tail -3 $out/launches_$tag.out;;
ncu)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"block_kernel|count_tile_kernel" -s 4 -c 4 -o $out/prof_$tag \
     python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $out/ncu_$tag.out 2>&1; tail -3 $out/ncu_$tag.out;;
bench_h64_minb)
  PG_BLOCK_MINB=1 timeout 1200 python bench.py --steps 3 --warmup 3 --workload h64s --no-cpu-baseline > $out/bench_h64s_minb1_$tag.json 2> $out/bench_h64s_minb1_$tag.err; tail -c 1200 $out/bench_h64s_minb1_$tag.json | head -c 1200; tail -3 $out/bench_h64s_minb1_$tag.err;;
ncu_h64)
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"skeleton_kernel|block_kernel" -s 3 -c 2 -o $out/prof_h64_$tag \
     python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload h64s > $out/ncu_h64_$tag.out 2>&1; tail -3 $out/ncu_h64_$tag.out;;
ncu_count_big)
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"count_tile_kernel" -s 14 -c 1 -o $out/prof_countbig_$tag \
     python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload cfg3s > $out/ncu_countbig_$tag.out 2>&1; tail -3 $out/ncu_countbig_$tag.out;;
ncu_count)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"count_tile_kernel" -s 3 -c 2 -o $out/prof_count_$tag \
     python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $out/ncu_count_$tag.out 2>&1; tail -3 $out/ncu_count_$tag.out;;
sanitizer)
  timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_hmm.py tests/test_gpu_counting.py -m gpu -q -p no:cacheprovider \
     -k "reference_vector or edge_cases or options or kmercounter_vectors" 2>&1 | tail -60 > $out/sanitizer_$tag.log; tail -8 $out/sanitizer_$tag.log;;
esac
done
