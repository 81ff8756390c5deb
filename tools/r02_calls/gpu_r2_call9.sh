#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_vit.py -m gpu -q -x -s -k "attention_tcgen05" > gpurun_out/c9_t_attn.log 2>&1; echo "attn test rc=$?"; grep -E "max \|err\||passed|failed|Error|error" gpurun_out/c9_t_attn.log | tail -15
timeout 120 python tools/prof_vit_attn.py --images 512 2>&1 | tail -2
timeout 120 python tools/prof_vit_attn.py --images 128 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "cta_pairs" > gpurun_out/c9_t_pair.log 2>&1; echo "pair test rc=$?"; tail -5 gpurun_out/c9_t_pair.log
timeout 600 python -m pytest tests/test_vit.py -m gpu -q -x > gpurun_out/c9_t_vit.log 2>&1; echo "vit test rc=$?"; tail -5 gpurun_out/c9_t_vit.log
timeout 400 python bench.py --config 5 --no-cpu-baseline > gpurun_out/c9_bench_vit.json 2> gpurun_out/c9_bench_vit.err; cat gpurun_out/c9_bench_vit.json | cut -c1-330; grep -o '"clocks.*' gpurun_out/c9_bench_vit.json | cut -c1-600
DFD_LIB_PATH=build/variants/libdfd_vitres0.so timeout 400 python bench.py --config 5 --no-cpu-baseline > gpurun_out/c9_bench_vit_res0.json 2> gpurun_out/c9_bench_vit_res0.err; cat gpurun_out/c9_bench_vit_res0.json | cut -c1-330
