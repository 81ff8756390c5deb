#!/bin/bash
# per-role wait accounting (DFD_GEMM_DBG=32 [+7 skeleton]) on a few shapes; every run bounded
for s in "32 16 12544 0 0 0" "240 40 784 0 0 1" "80 480 196 1 0 0" "480 80 196 0 1 1" "1152 192 49 0 1 1"; do set -- $s
  for d in 32 39; do
    echo "[dbg=$d]"; DFD_GEMM_DBG=$d timeout 60 python tools/prof_gemm.py --K $1 --N $2 --HW $3 --act $4 --gate $5 --res $6 --frames 1024 --iters 1 2>&1 | tail -32
  done
done
