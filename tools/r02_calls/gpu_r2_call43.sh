#!/bin/bash
# call 43: final kernels (cluster SE gate for <= 128 frames per call): full GPU suite, smoke, bench, launch list at the bench size, batch sweep
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
tag=r02i
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$tag.log
CMD="python tools/prof_step.py --videos 64 --frames 32 --iters 2"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 68 -c 68 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_list_$tag.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 68 -c 68 --csv --log-file gpurun_out/launches_f8_$tag.csv python tools/prof_step.py --videos 1 --frames 8 > gpurun_out/ncu_f8_$tag.log 2>&1; echo "f8 list rc=$?"
timeout 300 python tools/sweep_batch.py > gpurun_out/sweep_batch_$tag.log 2>&1; head -13 gpurun_out/sweep_batch_$tag.log
