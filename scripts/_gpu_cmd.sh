PG_HMM_B=16 timeout 900 python -m pytest tests/test_gpu_hmm.py -m gpu -q --timeout 600 -p no:cacheprovider -k "boundaries" 2>&1 | tail -25
timeout 900 python -m pytest tests/test_gpu_hmm.py tests/test_gpu_counting.py -m gpu -q --timeout 600 -p no:cacheprovider -k "boundaries or partitioned" 2>&1 | tail -8
