# 8-GPU lease: the north-star configuration (configs[3]: 5 M variants, 64 haplotypes, 30x = 189 GB of FASTQ), one sample sharded
nvidia-smi --query-gpu=index,name,memory.total --format=csv | head -9
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_n8_r2.json 2> gpurun_out/bench_n8_r2.err
tail -c 3000 gpurun_out/bench_n8_r2.json; tail -8 gpurun_out/bench_n8_r2.err | cut -c1-400
