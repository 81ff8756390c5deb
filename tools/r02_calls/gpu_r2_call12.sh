#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_vit.py -m gpu -q -x -k "cta_pairs or vit" > gpurun_out/c12_t.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/c12_t.log
timeout 300 python tools/prof_gemm_pair.py --images 512 --impls 3 2>&1 | tail -4
timeout 400 python bench.py --config 5 --no-cpu-baseline > gpurun_out/c12_bench_vit.json 2> gpurun_out/c12_bench_vit.err; cat gpurun_out/c12_bench_vit.json | cut -c1-330; grep -o '"clocks.*' gpurun_out/c12_bench_vit.json | cut -c1-700
CMDV="python tools/bench_vit.py --batch 512 --iters 1"
timeout 300 $CMDV > gpurun_out/c12_vit_plain.log 2>&1; tail -2 gpurun_out/c12_vit_plain.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_tensor.sum --clock-control none -s 176 -c 88 --csv --log-file gpurun_out/c12_launches_vit.csv $CMDV > gpurun_out/c12_ncu_list_vit.log 2>&1
echo "vit launch list rc=$?"
