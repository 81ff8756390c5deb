#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "cta_pairs" > gpurun_out/c8_t_pair.log 2>&1; echo "pair test rc=$?"; tail -15 gpurun_out/c8_t_pair.log
timeout 300 python tools/prof_gemm_pair.py --images 512 > gpurun_out/c8_prof_pair.log 2>&1; cat gpurun_out/c8_prof_pair.log | tail -10
timeout 600 python -m pytest tests/test_vit.py -m gpu -q -x > gpurun_out/c8_t_vit.log 2>&1; echo "vit test rc=$?"; tail -8 gpurun_out/c8_t_vit.log
timeout 400 python bench.py --config 5 --no-cpu-baseline > gpurun_out/c8_bench_vit.json 2> gpurun_out/c8_bench_vit.err; cat gpurun_out/c8_bench_vit.json | cut -c1-1600
