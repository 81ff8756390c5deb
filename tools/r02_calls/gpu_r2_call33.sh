#!/bin/bash
# call 33: which role bounds the stem row kernel — timing builds that skip the builders' work (1), the epilogue's (2), the MMAs (4), the raw loads (8)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in default stemdbg1 stemdbg2 stemdbg3 stemdbg4 stemdbg7 stemdbg8 stemdbg15; do
  if [ $v = default ]; then unset DFD_LIB_PATH; else export DFD_LIB_PATH=build/variants/libdfd_$v.so; fi
  timeout 120 python tools/time_classes.py --only stem --iters 3 2>&1 | tail -1
done
