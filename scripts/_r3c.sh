bash scripts/gpu_round.sh r3c pytest smoke
STEPS=4 bash scripts/gpu_round.sh r3c bench
