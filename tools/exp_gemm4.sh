#!/bin/bash
# all pointwise-GEMM shapes of the network at 1024 frames, bounded runs
for s in "32 16 12544 0 0 0" "96 24 3136 0 0 0" "144 24 3136 0 0 1" "144 40 784 0 0 0" "240 40 784 0 0 1" "16 96 12544 1 0 0" "24 144 3136 1 0 0" "40 240 784 1 0 0" "80 480 196 1 0 0" "112 672 196 1 0 0" "192 1152 49 1 0 0" "240 80 196 0 1 0" "480 80 196 0 1 1" "480 112 196 0 1 0" "672 112 196 0 1 1" "672 192 49 0 1 0" "1152 192 49 0 1 1" "1152 320 49 0 1 0"; do set -- $s
  for env in "X=0" $EXTRA_ENVS; do
    echo -n "[$env] "; env $env timeout 60 python tools/prof_gemm.py --K $1 --N $2 --HW $3 --act $4 --gate $5 --res $6 --frames 1024 --iters 3 2>&1 | tail -1
  done
done
