#!/bin/bash
# call 44: transformer warps of the gated project GEMMs (8 / 12 / 16)
set -u
export PYTHONUNBUFFERED=1
for v in default xw12 xw16; do
  if [ $v = default ]; then unset DFD_LIB_PATH; else export DFD_LIB_PATH=build/variants/libdfd_$v.so; fi
  echo "== $v"
  timeout 100 python tools/prof_gemm.py --K 1152 --N 192 --HW 49 --frames 2048 --gate 1 --res 1 --act 0 --iters 3 2>&1 | tail -1
  timeout 100 python tools/prof_gemm.py --K 1152 --N 320 --HW 49 --frames 2048 --gate 1 --res 0 --act 0 --iters 3 2>&1 | tail -1
  timeout 100 python tools/prof_gemm.py --K 672 --N 112 --HW 196 --frames 2048 --gate 1 --res 1 --act 0 --iters 3 2>&1 | tail -1
  timeout 100 python tools/prof_gemm.py --K 480 --N 80 --HW 196 --frames 2048 --gate 1 --res 1 --act 0 --iters 3 2>&1 | tail -1
  timeout 100 python tools/prof_gemm.py --K 672 --N 192 --HW 49 --frames 2048 --gate 1 --res 0 --act 0 --iters 3 2>&1 | tail -1
done
