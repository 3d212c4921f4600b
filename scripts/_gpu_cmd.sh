bash scripts/gpu_round.sh s2d launches
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"count_tile_kernel|basis_kernel|scan_kernel|block_kernel" -s 8 -c 6 -o gpurun_out/prof_s2d python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_s2d.out 2>&1; tail -3 gpurun_out/ncu_s2d.out
