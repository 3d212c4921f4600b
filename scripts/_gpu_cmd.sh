timeout 900 python -m pytest tests/test_gpu_counting.py -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -25
