nvidia-smi --query-gpu=index,name --format=csv | head -4
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_golden_reference_fixture.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/pytest_multi_r2y.log; cat gpurun_out/pytest_multi_r2y.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2_r2y.json 2> gpurun_out/bench_n2_r2y.err
tail -c 1500 gpurun_out/bench_n2_r2y.json; tail -4 gpurun_out/bench_n2_r2y.err | cut -c1-300
