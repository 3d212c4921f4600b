# 2-GPU lease: NCCL parity test, sharded C++ host, sharded bench (configs[2] over two ranks)
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_golden_reference_fixture.py -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/pytest_multi_r2j.log; tail -5 gpurun_out/pytest_multi_r2j.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2_r2j.json 2> gpurun_out/bench_n2_r2j.err
tail -c 2500 gpurun_out/bench_n2_r2j.json; tail -5 gpurun_out/bench_n2_r2j.err
