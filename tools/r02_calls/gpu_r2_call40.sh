#!/bin/bash
# call 40-41: TMA-store epilogue for N = 240, 480 (16-column boxes alone; then 32-column boxes + a 16-column tail): parity and A/B against direct stores
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_path.py -m gpu -q -x -k "gemm or golden or determinism" > gpurun_out/c40_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c40_pytest.log
for v in default notstore; do
  if [ $v = default ]; then unset DFD_LIB_PATH; else export DFD_LIB_PATH=build/variants/libdfd_$v.so; fi
  echo "== $v"
  timeout 100 python tools/prof_gemm.py --K 40 --N 240 --HW 784 --frames 2048 --gate 0 --res 0 --act 1 --iters 3 2>&1 | tail -1
  timeout 100 python tools/prof_gemm.py --K 80 --N 480 --HW 196 --frames 2048 --gate 0 --res 0 --act 1 --iters 3 2>&1 | tail -1
done
unset DFD_LIB_PATH
timeout 120 python tools/time_classes.py --iters 3 2>&1 | tail -1
