bash scripts/gpu_round.sh r2g pytest_new
for kb in 98304 40960; do
  PG_COUNT_PART_KB=$kb PG_BENCH_E2E_STEPS=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_occ${kb}_r2g.json 2> gpurun_out/bench_occ${kb}_r2g.err
  python - <<EOF
import json
d=json.loads([l for l in open("gpurun_out/bench_occ${kb}_r2g.json") if l.startswith("{")][0]); u=d["roofline"]["update_pass"]
print("slice_kb", $kb, "step", round(d["ms_per_step"],1), "update", round(u["ms"],1), "probe", round(u["probe_ms"],1), "scatter", round(u["ms"]-u["probe_ms"],1))
EOF
done
PG_COUNT_PART_KB=98304 bash scripts/gpu_round.sh r2g ncu_scatter
