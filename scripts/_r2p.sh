# re-entry check of HEAD: GPU tests, default bench line (full configs[2]), cluster checkpoint walk sweep, multi-sample hmm
bash scripts/gpu_round.sh r2p pytest
STEPS=4 bash scripts/gpu_round.sh r2p bench
bash scripts/gpu_round.sh r2p hmm_cluster
timeout 600 python scripts/bench_hmm.py --haplotypes 32 64 --variants 400000 --repeat 2 --samples 4 > gpurun_out/bench_hmm_s4_r2p.jsonl 2> gpurun_out/bench_hmm_s4_r2p.err; cat gpurun_out/bench_hmm_s4_r2p.jsonl
