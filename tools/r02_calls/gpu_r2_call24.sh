#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_vit.py -m gpu -q -x -k "cta_pairs or vit" > gpurun_out/c24_t.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/c24_t.log
timeout 300 python tools/prof_gemm_pair.py --images 512 --impls 3 2>&1 | tail -4
timeout 400 python bench.py --config 5 --no-cpu-baseline > gpurun_out/c24_bench_vit.json 2> gpurun_out/c24_bench_vit.err; cut -c1-330 gpurun_out/c24_bench_vit.json; grep -o '"clocks.*' gpurun_out/c24_bench_vit.json | cut -c1-600
