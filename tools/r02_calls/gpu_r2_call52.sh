#!/bin/bash
# call 52: ncu --set full of the GEMM kernels of one forward (512 frames) on the final code (12 transformer warps in the gated layers)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CMD2="python tools/prof_step.py --videos 16 --frames 32 --iters 2"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 33 -c 33 -f -o /tmp/full_gemm $CMD2 > gpurun_out/ncu_full_gemm_r02p.log 2>&1; echo "full rc=$?"
ncu -i /tmp/full_gemm.ncu-rep --page raw --csv > gpurun_out/full_gemm_r02p_raw.csv 2>/dev/null; wc -c gpurun_out/full_gemm_r02p_raw.csv
