#!/bin/bash
# call 51: final tree (comment / macro-guard edits only since call 47): full GPU suite, smoke, launch list for the traffic record, bench line
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
tag=r02p
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$tag.log
CMD="python tools/prof_step.py --videos 64 --frames 32 --iters 2"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 68 -c 68 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_list_$tag.log 2>&1; echo "launch list rc=$?"
