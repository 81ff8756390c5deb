#!/bin/bash
# call 32: epilogue / builder set counts of the stem row kernel after its instruction diet
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in default stem33 stem34 stem43 stem42 stem23 stem32 default; do
  if [ $v = default ]; then unset DFD_LIB_PATH; else export DFD_LIB_PATH=build/variants/libdfd_$v.so; fi
  timeout 120 python tools/time_classes.py --only stem 2>&1 | tail -1
done
