#!/bin/bash
# call 35: stem row kernel (two rows per tile, 16 raw stages): epilogue / builder set counts, role ablation
set -u
export PYTHONUNBUFFERED=1
for v in default stem33 stem34 stem43 stem42 stem23 stemdbg1 stemdbg2 stemdbg3 default; do
  if [ $v = default ]; then unset DFD_LIB_PATH; else export DFD_LIB_PATH=build/variants/libdfd_$v.so; fi
  timeout 120 python tools/time_classes.py --only stem --iters 3 2>&1 | tail -1
done
