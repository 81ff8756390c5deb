#!/bin/bash
# bench.py on N GPUs of one box, launched as the driver does.  Usage: bash tools/gpu_multi.sh <N> <tag>
N=$1; tag=$2
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.err; echo "bench n$N rc=$?"
cut -c1-400 gpurun_out/bench_n${N}_$tag.json; grep -o '"strong".*' gpurun_out/bench_n${N}_$tag.json | cut -c1-900
nvidia-smi topo -m > gpurun_out/topo_n${N}_$tag.txt 2>&1; grep -o "host_affinity[^,]*,[^,]*" gpurun_out/bench_n${N}_$tag.json | head -2; grep -o "\"e2e\": {[^}]*}" gpurun_out/bench_n${N}_$tag.json
