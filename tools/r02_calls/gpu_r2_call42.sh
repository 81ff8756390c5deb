#!/bin/bash
# call 42: SE gate of the wide layers on thread-block clusters (channels split over 8 CTAs): parity, A/B at 2048 and 8 frames
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_path.py -m gpu -q -x -k "se_gate or golden or determinism or graph" > gpurun_out/c42_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c42_pytest.log
for v in default nocluster; do
  if [ $v = default ]; then unset DFD_LIB_PATH; else export DFD_LIB_PATH=build/variants/libdfd_$v.so; fi
  echo "== $v"
  for fr in 2048 8; do
    timeout 60 python tools/prof_se.py --C 1152 --rd 48 --nparts 1 --frames $fr 2>&1 | tail -1
    timeout 60 python tools/prof_se.py --C 672 --rd 28 --nparts 2 --frames $fr 2>&1 | tail -1
    timeout 60 python tools/prof_se.py --C 480 --rd 20 --nparts 2 --frames $fr 2>&1 | tail -1
  done
  timeout 120 python tools/time_classes.py --iters 3 --only se_gate 2>&1 | tail -1
  timeout 120 python tools/time_classes.py --iters 5 --videos 1 --frames 8 --only se_gate 2>&1 | tail -1
done
unset DFD_LIB_PATH
timeout 200 python tools/sweep_batch.py 2>&1 | head -9
