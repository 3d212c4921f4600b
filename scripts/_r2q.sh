# row-per-lane checkpoint walk: correctness under both tiles, then timing at H = 32 / 64
for m in 1 2; do
  PG_SKELETON_TILE=$m timeout 600 python -m pytest tests/test_gpu_hmm.py tests/test_gpu_at_size.py -m gpu -q -x -p no:cacheprovider -k "not cluster" 2>&1 | tail -3
done
for m in 0 1 2; do
  PG_SKELETON_TILE=$m timeout 600 python scripts/bench_hmm.py --haplotypes 32 64 --variants 400000 --repeat 2 > gpurun_out/bench_hmm_tile${m}_r2q.jsonl 2> gpurun_out/bench_hmm_tile${m}_r2q.err
  python - <<PY
import json
for l in open("gpurun_out/bench_hmm_tile${m}_r2q.jsonl"):
    d=json.loads(l); print("tile", $m, "H", d["haplotypes"], "skeleton_ms", round(d["skeleton_ms"],2), "blocks_ms", round(d["blocks_ms"],2))
PY
  tail -2 gpurun_out/bench_hmm_tile${m}_r2q.err
done
