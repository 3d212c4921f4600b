timeout 900 python -m pytest tests/test_gpu_hmm.py -m gpu -q -x -p no:cacheprovider -k "cluster or boundaries or multi_chromosome" 2>&1 | tail -6
bash scripts/gpu_round.sh r2l hmm_cluster
