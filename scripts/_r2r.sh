scripts/microbench/latency > gpurun_out/latency_r2r.txt 2>&1; cat gpurun_out/latency_r2r.txt
PG_SKELETON_TILE=1 timeout 300 python -m pytest tests/test_gpu_hmm.py -m gpu -q -x -p no:cacheprovider -k "test_hmm_matches_oracle" 2>&1 | tail -40 > gpurun_out/tile1_fail_r2r.log
for m in 0 1; do
PG_SKELETON_TILE=$m timeout 600 ncu --set full --clock-control none --import-source on -k regex:"skeleton" -c 1 -o gpurun_out/prof_skel_tile${m}_r2r \
   python scripts/bench_hmm.py --haplotypes 32 --variants 100000 --repeat 1 > gpurun_out/ncu_skel_tile${m}_r2r.out 2>&1; tail -2 gpurun_out/ncu_skel_tile${m}_r2r.out
done
