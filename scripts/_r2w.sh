PG_SKELETON_TILE=4 timeout 900 python -m pytest tests/test_gpu_hmm.py tests/test_gpu_at_size.py -m gpu -q -p no:cacheprovider -k "not cluster" 2>&1 | tail -4
for m in 3 4; do
  PG_SKELETON_TILE=$m timeout 600 python scripts/bench_hmm.py --haplotypes 32 64 --variants 400000 --repeat 2 > gpurun_out/bench_hmm_tile${m}_r2w.jsonl 2> gpurun_out/bench_hmm_tile${m}_r2w.err
  python - <<PY
import json
for l in open("gpurun_out/bench_hmm_tile${m}_r2w.jsonl"):
    d=json.loads(l); print("tile", $m, "H", d["haplotypes"], "skeleton_ms", round(d["skeleton_ms"],2), "blocks_ms", round(d["blocks_ms"],2))
PY
  tail -2 gpurun_out/bench_hmm_tile${m}_r2w.err
done
