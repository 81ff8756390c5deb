#!/bin/bash
set -u
export PYTHONUNBUFFERED=1
S="--K 576 --N 64 --HW 3136 --frames 96 --gate 0 --res 0 --act 3"
for d in 3 259 515 771 775; do echo "== dbg $d"; DFD_GEMM_DBG=$d python tools/prof_gemm.py $S | tail -1; done
echo "== role waits dbg 32+3"; DFD_GEMM_DBG=35 python tools/prof_gemm.py $S --iters 1 | grep -E "warp  [089] |warp 1[0-2] " | head -12
