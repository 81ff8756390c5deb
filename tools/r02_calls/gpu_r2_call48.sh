#!/bin/bash
# call 48: same-box A/B of 12 vs 8 transformer warps on the whole step, then the bench line once more (box-to-box spread)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in default xw8 default xw8; do
  if [ $v = default ]; then unset DFD_LIB_PATH; else export DFD_LIB_PATH=build/variants/libdfd_$v.so; fi
  timeout 120 python tools/time_classes.py --iters 5 2>&1 | tail -1
done
unset DFD_LIB_PATH
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02k.json 2> gpurun_out/bench_r02k.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_r02k.json
nvidia-smi --query-gpu=name,temperature.gpu,power.draw,clocks.sm,clocks.mem --format=csv
