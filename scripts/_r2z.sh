timeout 600 ncu --set full --clock-control none --import-source on -k regex:"block_kernel" -c 1 -o gpurun_out/prof_block_h32_r2z \
   python scripts/bench_hmm.py --haplotypes 32 --variants 200000 --repeat 1 > gpurun_out/ncu_block_r2z.out 2>&1; tail -2 gpurun_out/ncu_block_r2z.out
