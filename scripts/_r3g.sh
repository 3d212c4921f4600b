for mb in 1024 2560; do
  PG_COUNT_SUPER_MB=$mb timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_super${mb}_r3g.json 2> gpurun_out/bench_super${mb}_r3g.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_super${mb}_r3g.json") if l.startswith("{")][0]); u=d["roofline"]["update_pass"]
print("super_mb", $mb, "step", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["ms_per_step"],1), "update", round(u["ms"],1), "probe", round(u["probe_ms"],1), "passes", u["probe_passes"])
PY
done
