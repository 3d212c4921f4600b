bash scripts/gpu_round.sh r1f pytest smoke
timeout 900 python bench.py > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; tail -c 600 gpurun_out/bench_r1f.json; tail -3 gpurun_out/bench_r1f.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1f.json 2> gpurun_out/bench_ref_r1f.err; tail -c 900 gpurun_out/bench_ref_r1f.json; tail -3 gpurun_out/bench_ref_r1f.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/launches_r1f.out 2>&1; tail -2 gpurun_out/launches_r1f.out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"count_tile_kernel|basis_kernel|^scan_kernel|block_kernel" -s 10 -c 7 -o gpurun_out/prof_r1f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r1f.out 2>&1; tail -2 gpurun_out/ncu_r1f.out
