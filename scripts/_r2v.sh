PG_SKELETON_TILE=3 timeout 900 python -m pytest tests/test_gpu_hmm.py tests/test_gpu_at_size.py -m gpu -q -p no:cacheprovider -k "not cluster" 2>&1 | tail -8
for m in 0 3; do
  PG_SKELETON_TILE=$m timeout 600 python scripts/bench_hmm.py --haplotypes 32 64 --variants 400000 --repeat 2 > gpurun_out/bench_hmm_tile${m}_r2v.jsonl 2> gpurun_out/bench_hmm_tile${m}_r2v.err
  python - <<PY
import json
for l in open("gpurun_out/bench_hmm_tile${m}_r2v.jsonl"):
    d=json.loads(l); print("tile", $m, "H", d["haplotypes"], "skeleton_ms", round(d["skeleton_ms"],2), "blocks_ms", round(d["blocks_ms"],2))
PY
  tail -2 gpurun_out/bench_hmm_tile${m}_r2v.err
done
PG_SKELETON_TILE=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"skeleton" -c 1 -o gpurun_out/prof_skel_lean_r2v \
   python scripts/bench_hmm.py --haplotypes 32 --variants 100000 --repeat 1 > gpurun_out/ncu_skel_lean_r2v.out 2>&1; tail -2 gpurun_out/ncu_skel_lean_r2v.out
