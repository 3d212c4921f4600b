bash scripts/gpu_round.sh r2i pytest
STEPS=3 bash scripts/gpu_round.sh r2i bench
