#!/bin/bash
# gated project shapes under env variants
for s in "240 80 196 0 1 0" "480 80 196 0 1 1" "480 112 196 0 1 0" "672 112 196 0 1 1" "672 192 49 0 1 0" "1152 192 49 0 1 1" "1152 320 49 0 1 0"; do set -- $s
  for env in "X=0" $EXTRA_ENVS; do
    echo -n "[$env] "; env $env timeout 60 python tools/prof_gemm.py --K $1 --N $2 --HW $3 --act $4 --gate $5 --res $6 --frames 2048 --iters 3 2>&1 | tail -1
  done
done
