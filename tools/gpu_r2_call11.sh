#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "cta_pairs" > gpurun_out/c11_t_pair.log 2>&1; echo "pair test rc=$?"; tail -3 gpurun_out/c11_t_pair.log
echo "== warp store (default)"; timeout 300 python tools/prof_gemm_pair.py --images 512 --impls 3 2>&1 | tail -4
for v in pair_group pair_group_nostore pair_warp_nostore; do echo "== $v"; DFD_LIB_PATH=build/variants/libdfd_$v.so timeout 300 python tools/prof_gemm_pair.py --images 512 --impls 3 2>&1 | tail -4; done
