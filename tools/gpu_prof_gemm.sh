#!/bin/bash
# ncu source-level profile of one GEMM shape; exports CSV pages on the box.  Usage: bash tools/gpu_prof_gemm.sh <tag> <prof_gemm args...>
tag=$1; shift
mkdir -p gpurun_out
python tools/prof_gemm.py "$@" > gpurun_out/gemm_plain_$tag.log 2>&1 || { tail -5 gpurun_out/gemm_plain_$tag.log; exit 1; }
cat gpurun_out/gemm_plain_$tag.log
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o /tmp/g_$tag python tools/prof_gemm.py "$@" --iters 2 > gpurun_out/ncu_gemm_$tag.log 2>&1
ncu -i /tmp/g_$tag.ncu-rep --page raw --csv > gpurun_out/gemm_${tag}_raw.csv 2>/dev/null
ncu -i /tmp/g_$tag.ncu-rep --page source --csv > gpurun_out/gemm_${tag}_source.csv 2>/dev/null
ls -la /tmp/g_$tag.ncu-rep gpurun_out/gemm_${tag}_*
