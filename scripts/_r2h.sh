bash scripts/gpu_round.sh r2h pytest_new
PG_BENCH_E2E_STEPS=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_probe3_r2h.json 2> gpurun_out/bench_probe3_r2h.err
python - <<EOF
import json
d=json.loads([l for l in open("gpurun_out/bench_probe3_r2h.json") if l.startswith("{")][0]); u=d["roofline"]["update_pass"]
print("step", round(d["ms_per_step"],1), "update", round(u["ms"],1), "probe", round(u["probe_ms"],1), "scatter", round(u["ms"]-u["probe_ms"],1), "parity", d["parity"]["ok"])
EOF
bash scripts/gpu_round.sh r2h ncu_probe
