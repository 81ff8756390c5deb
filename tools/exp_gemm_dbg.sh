#!/bin/bash
for s in "32 16 12544 0 0 0" "240 40 784 0 0 0" "480 80 196 0 1 0" "1152 192 49 0 1 0" "80 480 196 1 0 0"; do set -- $s
  for d in 0 1 2 3 4 7; do
    echo -n "[dbg=$d] "; DFD_GEMM_DBG=$d python tools/prof_gemm.py --K $1 --N $2 --HW $3 --act $4 --gate $5 --res $6 --frames 1024 --iters 3
  done
done
