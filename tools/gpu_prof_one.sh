#!/bin/bash
# ncu source-level profile of one kernel-level driver; exports CSV pages on the box.
# Usage: bash tools/gpu_prof_one.sh <tag> <kernel-regex> <script.py> <args...>
tag=$1; regex=$2; script=$3; shift 3
mkdir -p gpurun_out
python $script "$@" > gpurun_out/one_plain_$tag.log 2>&1 || { tail -5 gpurun_out/one_plain_$tag.log; exit 1; }
cat gpurun_out/one_plain_$tag.log
ncu --set full --clock-control none --import-source on -k regex:$regex -s 2 -c 1 -f -o /tmp/o_$tag python $script "$@" --iters 2 > gpurun_out/ncu_one_$tag.log 2>&1
ncu -i /tmp/o_$tag.ncu-rep --page raw --csv > gpurun_out/one_${tag}_raw.csv 2>/dev/null
ncu -i /tmp/o_$tag.ncu-rep --page source --csv > gpurun_out/one_${tag}_source.csv 2>/dev/null
ls -la gpurun_out/one_${tag}_*
