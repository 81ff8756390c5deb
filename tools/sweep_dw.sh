#!/bin/bash
# sweep depthwise kernel choices over the 12 network shapes: window-per-row kernel vs row-marching kernel (CTA size limits)
for s in "32 3 1 112" "96 3 2 112" "144 3 1 56" "144 5 2 56" "240 5 1 28" "240 3 2 28" "480 3 1 14" "480 5 1 14" "672 5 1 14" "672 5 2 14" "1152 5 1 7" "1152 3 1 7"; do set -- $s
  echo "== C=$1 k=$2 s=$3 H=$4"
  DFD_DW_MARCH=0 python tools/prof_dw.py --C $1 --k $2 --s $3 --H $4 --frames ${FRAMES:-1024} --iters 3 | sed 's/^.*best/  window          best/'
  for mt in ${MAXTS:-128 192 256}; do
    DFD_DW_MARCH=1 DFD_DW_MAXT=$mt python tools/prof_dw.py --C $1 --k $2 --s $3 --H $4 --frames ${FRAMES:-1024} --iters 3 | sed "s/^.*best/  march maxt=$mt  best/"
  done
done
