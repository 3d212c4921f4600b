timeout 900 python -m pytest tests/test_gpu_hmm.py tests/test_gpu_pipeline.py -m gpu -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_hmm.log; tail -5 gpurun_out/pytest_hmm.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; python -c "
import json; d=json.load(open('gpurun_out/bench_b.json')); print(d['value'], d['clocks']['ms_per_step_without_sampler'], d['e2e']['ms_per_step'], d['roofline']['stage_ms'])"; tail -3 gpurun_out/bench_b.err
