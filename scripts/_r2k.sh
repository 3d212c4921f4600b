bash scripts/gpu_round.sh r2k pytest_new
timeout 600 python -m pytest tests/test_gpu_counting.py tests/test_gpu_pipeline.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
PG_BENCH_E2E_STEPS=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_scat3_r2k.json 2> gpurun_out/bench_scat3_r2k.err
python - <<EOF
import json
d=json.loads([l for l in open("gpurun_out/bench_scat3_r2k.json") if l.startswith("{")][0]); u=d["roofline"]["update_pass"]
print("step", round(d["ms_per_step"],1), "update", round(u["ms"],1), "probe", round(u["probe_ms"],1), "scatter", round(u["ms"]-u["probe_ms"],1), "parity", d["parity"]["ok"])
EOF
bash scripts/gpu_round.sh r2k ncu_scatter
