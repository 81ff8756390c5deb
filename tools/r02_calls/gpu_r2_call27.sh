#!/bin/bash
set -u
export PYTHONUNBUFFERED=1
for v in "" waitlane0; do
  lib=${v:+build/variants/libdfd_$v.so}; echo "== ${v:-all lanes poll}"
  DFD_LIB_PATH=$lib python tools/prof_gemm.py --K 80 --N 480 --HW 196 --frames 2048 --gate 0 --res 0 --act 1 --iters 10 | tail -1
  DFD_LIB_PATH=$lib python tools/prof_gemm.py --K 40 --N 240 --HW 784 --frames 2048 --gate 0 --res 0 --act 1 --iters 10 | tail -1
  DFD_LIB_PATH=$lib python tools/prof_gemm.py --K 1152 --N 192 --HW 49 --frames 2048 --gate 1 --res 1 --act 0 --iters 10 | tail -1
  DFD_LIB_PATH=$lib python tools/prof_gemm.py --K 672 --N 112 --HW 196 --frames 2048 --gate 1 --res 1 --act 0 --iters 10 | tail -1
  DFD_LIB_PATH=$lib timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items()})"
done
