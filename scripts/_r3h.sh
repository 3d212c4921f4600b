bash scripts/gpu_round.sh r3h pytest smoke
STEPS=4 bash scripts/gpu_round.sh r3h bench > /dev/null
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_r3h.json") if l.startswith("{")][0]); u=d["roofline"]["update_pass"]
print("step", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["ms_per_step"],1), "update", round(u["ms"],1), "probe", round(u["probe_ms"],1), "passes", u["probe_passes"], "parity", d["parity"]["ok"], "cpu", round(d["cpu_baseline"]["value"]))
PY
