timeout 600 python -m pytest tests/test_gpu_hmm.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
PG_SKELETON_WIDE=1 timeout 600 python -m pytest tests/test_gpu_hmm.py tests/test_gpu_at_size.py -m gpu -q -p no:cacheprovider -k "lean or boundaries or at_size or h64 or P65 or 65" 2>&1 | tail -3
for m in 0 1; do
  PG_SKELETON_WIDE=$m timeout 600 python scripts/bench_hmm.py --haplotypes 32 64 --variants 400000 --repeat 2 > gpurun_out/bench_hmm_wide${m}_r3f.jsonl 2> gpurun_out/bench_hmm_wide${m}_r3f.err
  python - <<PY
import json
for l in open("gpurun_out/bench_hmm_wide${m}_r3f.jsonl"):
    d=json.loads(l); print("wide", $m, "H", d["haplotypes"], "skeleton_ms", round(d["skeleton_ms"],2), "blocks_ms", round(d["blocks_ms"],3))
PY
  tail -2 gpurun_out/bench_hmm_wide${m}_r3f.err
done
