timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_all.log; tail -5 gpurun_out/pytest_all.log
PG_TRACE=1 timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; python -c "
import json; d=json.load(open('gpurun_out/bench_b.json')); print(d['value'], d['ms_per_step'], d['clocks'], d['e2e'], d['roofline']['stage_ms'], d['roofline']['frac'])"; grep genotype_run gpurun_out/bench_b.err | tail -14
