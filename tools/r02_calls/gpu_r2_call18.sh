#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_vit.py -m gpu -q -x -k "cta_pairs or vit" > gpurun_out/c18_t.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/c18_t.log
timeout 400 python bench.py --config 5 --no-cpu-baseline > gpurun_out/c18_bench_vit.json 2> gpurun_out/c18_bench_vit.err; cat gpurun_out/c18_bench_vit.json | cut -c1-330; grep -o '"clocks.*' gpurun_out/c18_bench_vit.json | cut -c1-700
CMDV="python tools/bench_vit.py --batch 512 --iters 1"
timeout 300 $CMDV > gpurun_out/c18_vit_plain.log 2>&1; tail -1 gpurun_out/c18_vit_plain.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_tensor.sum --clock-control none -s 176 -c 88 --csv --log-file gpurun_out/c18_launches_vit.csv $CMDV > gpurun_out/c18_ncu_list_vit.log 2>&1
echo "vit launch list rc=$?"
