nvidia-smi --query-gpu=index,name,memory.total --format=csv | head -6; free -g | head -2; nproc
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 2 --warmup 3 > gpurun_out/bench_n4_r3d.json 2> gpurun_out/bench_n4_r3d.err
echo "rc=$?"; tail -c 1200 gpurun_out/bench_n4_r3d.json; tail -6 gpurun_out/bench_n4_r3d.err | cut -c1-300
