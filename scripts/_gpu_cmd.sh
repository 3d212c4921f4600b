for d in 2 1 3; do
timeout 600 python bench.py --inflight $d --no-cpu-baseline > gpurun_out/bench_if$d.json 2> gpurun_out/bench_if$d.err; python -c "
import json; d=json.load(open('gpurun_out/bench_if$d.json')); print($d, 'value', d['value'], d['ms_per_step'], 'serial', d['serial_ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['serial_ms_per_step'], d['clocks']['samples'])"; tail -3 gpurun_out/bench_if$d.err
done
