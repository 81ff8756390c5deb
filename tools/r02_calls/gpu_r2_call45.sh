#!/bin/bash
# call 45-46: 12 transformer warps in 4 / 3 / 2 groups (after fixing the row passes of a group of 96 threads); then 16 / 12 transformer warps with 4 epilogue warps: parity + timing
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in xw16e4 xw12e4; do
  export DFD_LIB_PATH=build/variants/libdfd_$v.so
  echo "== $v"
  timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm_tcgen05" 2>&1 | tail -1
  timeout 100 python tools/prof_gemm.py --K 1152 --N 192 --HW 49 --frames 2048 --gate 1 --res 1 --act 0 --iters 3 2>&1 | tail -1
  timeout 100 python tools/prof_gemm.py --K 1152 --N 320 --HW 49 --frames 2048 --gate 1 --res 0 --act 0 --iters 3 2>&1 | tail -1
  timeout 100 python tools/prof_gemm.py --K 672 --N 112 --HW 196 --frames 2048 --gate 1 --res 1 --act 0 --iters 3 2>&1 | tail -1
  timeout 100 python tools/prof_gemm.py --K 480 --N 80 --HW 196 --frames 2048 --gate 1 --res 1 --act 0 --iters 3 2>&1 | tail -1
  timeout 100 python tools/prof_gemm.py --K 672 --N 192 --HW 49 --frames 2048 --gate 1 --res 0 --act 0 --iters 3 2>&1 | tail -1
done
