bash scripts/gpu_round.sh r2o pytest_new
timeout 600 python -m pytest tests/test_gpu_counting.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
PG_COUNT_DEBUG=0 python scripts/count_dissect.py 0
PG_BENCH_E2E_STEPS=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_scat4_r2o.json 2> gpurun_out/bench_scat4_r2o.err
python - <<EOF
import json
d=json.loads([l for l in open("gpurun_out/bench_scat4_r2o.json") if l.startswith("{")][0]); u=d["roofline"]["update_pass"]
print("step", round(d["ms_per_step"],1), "update", round(u["ms"],1), "probe", round(u["probe_ms"],1), "scatter", round(u["ms"]-u["probe_ms"],1), "parity", d["parity"]["ok"])
EOF
