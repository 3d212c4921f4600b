timeout 900 python -m pytest tests/test_gpu_counting.py tests/test_gpu_at_size.py tests/test_gpu_pipeline.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
STEPS=4 bash scripts/gpu_round.sh r3e bench | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); u=d['roofline']['update_pass']
        print('step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['ms_per_step'],1), 'update', round(u['ms'],1), 'probe', round(u['probe_ms'],1), 'passes', u['probe_passes'], 'parity', d['parity']['ok'])
"
