timeout 900 python -m pytest tests/test_gpu_hmm.py tests/test_gpu_pipeline.py -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_hmm.log; tail -15 gpurun_out/pytest_hmm.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; python -c "
import json; d=json.load(open('gpurun_out/bench_a.json')); print(d['value'], d['ms_per_step'], d['clocks'], d['e2e'], d['roofline']['stage_ms'])"; tail -5 gpurun_out/bench_a.err
