timeout 900 python -m pytest tests/test_gpu_hmm.py tests/test_gpu_at_size.py tests/test_gpu_pipeline.py tests/test_demo_config0.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8
for m in 0 1; do
  PG_BLOCK_LEAN=$m timeout 600 python scripts/bench_hmm.py --haplotypes 32 64 --variants 400000 --repeat 3 > gpurun_out/bench_hmm_blean${m}_r3b.jsonl 2> gpurun_out/bench_hmm_blean${m}_r3b.err
  python - <<PY
import json
for l in open("gpurun_out/bench_hmm_blean${m}_r3b.jsonl"):
    d=json.loads(l); print("block_lean", $m, "H", d["haplotypes"], "skeleton_ms", round(d["skeleton_ms"],2), "blocks_ms", round(d["blocks_ms"],3), "frac", round(d["frac"],3))
PY
  tail -2 gpurun_out/bench_hmm_blean${m}_r3b.err
done
