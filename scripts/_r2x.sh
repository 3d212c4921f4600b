bash scripts/gpu_round.sh r2x pytest
STEPS=4 bash scripts/gpu_round.sh r2x bench
bash scripts/gpu_round.sh r2x bench_index
bash scripts/gpu_round.sh r2x sanitizer
