for dbg in 0 16 32 64 96; do
  PG_COUNT_DEBUG=$dbg PG_BENCH_E2E_STEPS=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_dbg${dbg}_r2m.json 2> gpurun_out/bench_dbg${dbg}_r2m.err
  python - <<EOF
import json
d=json.loads([l for l in open("gpurun_out/bench_dbg${dbg}_r2m.json") if l.startswith("{")][0]); u=d["roofline"]["update_pass"]
print("debug", $dbg, "update", round(u["ms"],1), "probe", round(u["probe_ms"],1), "scatter", round(u["ms"]-u["probe_ms"],1))
EOF
done
