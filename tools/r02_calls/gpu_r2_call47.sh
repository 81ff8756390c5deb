#!/bin/bash
# call 47: evidence on the final kernels (12 transformer warps in the gated GEMMs): full GPU suite, smoke, bench, launch list
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
tag=r02j
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$tag.log
CMD="python tools/prof_step.py --videos 64 --frames 32 --iters 2"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 68 -c 68 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_list_$tag.log 2>&1; echo "launch list rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_$tag.json
