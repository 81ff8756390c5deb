#!/bin/bash
# call 49: N = 480 expand layers as 3 chunks of 160 columns (TMA-store epilogue, 64-byte-aligned rows) instead of 2 x 240 (direct stores)
set -u
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm_tcgen05" 2>&1 | tail -1
for v in default nochunk32 default nochunk32; do
  if [ $v = default ]; then unset DFD_LIB_PATH; else export DFD_LIB_PATH=build/variants/libdfd_$v.so; fi
  timeout 100 python tools/prof_gemm.py --K 80 --N 480 --HW 196 --frames 2048 --gate 0 --res 0 --act 1 --iters 5 2>&1 | tail -1
done
