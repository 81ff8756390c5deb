#!/bin/bash
# call 38: what bounds the 7x7 expand GEMM (192 -> 1152) and the gated 7x7 project GEMM (1152 -> 192): role ablation with DFD_GEMM_DBG
set -u
export PYTHONUNBUFFERED=1
echo "== expand 192->1152 @7x7"
for d in 0 4 2 1 6 32; do echo "dbg $d"; DFD_GEMM_DBG=$d timeout 100 python tools/prof_gemm.py --K 192 --N 1152 --HW 49 --frames 2048 --gate 0 --res 0 --act 1 --iters 3 2>&1 | tail -24; done
echo "== project 1152->192 @7x7 gated + residual"
for d in 0 64 4 2 1 32; do echo "dbg $d"; DFD_GEMM_DBG=$d timeout 100 python tools/prof_gemm.py --K 1152 --N 192 --HW 49 --frames 2048 --gate 1 --res 1 --act 0 --iters 3 2>&1 | tail -30; done
