#!/bin/bash
# sweep depthwise kernel variants (build/variants/lib_<TW>_<MINB>.so) over the 12 network shapes
for lib in build/variants/lib_*.so; do
  echo "== $lib"
  for s in "32 3 1 112" "96 3 2 112" "144 3 1 56" "144 5 2 56" "240 5 1 28" "240 3 2 28" "480 3 1 14" "480 5 1 14" "672 5 1 14" "672 5 2 14" "1152 5 1 7" "1152 3 1 7"; do set -- $s
    DFD_LIB_PATH=$PWD/$lib python tools/prof_dw.py --C $1 --k $2 --s $3 --H $4 --frames 1024 --iters 3 | sed 's/frames=1024: best//'
  done
done
