#!/bin/bash
for lib in build/variants/lib_*.so; do
  echo "== $lib"
  for s in "16 96 12544 0 0 1" "24 144 3136 0 0 1" "40 240 784 0 0 1" "80 480 196 0 0 1" "112 672 196 0 0 1" "192 1152 49 0 0 1"; do set -- $s
    DFD_LIB_PATH=$PWD/$lib python tools/prof_gemm.py --K $1 --N $2 --HW $3 --gate $4 --res $5 --act $6 --frames 1024 --iters 3
  done
done
