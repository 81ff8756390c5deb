#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dw_small -s 1 -c 1 -f -o /tmp/dws7 python tools/prof_dw.py --C 1152 --k 5 --s 1 --H 7 --frames 2048 --iters 2 > gpurun_out/c15_ncu_dws7.log 2>&1; echo "rc=$?"
ncu -i /tmp/dws7.ncu-rep --page raw --csv > gpurun_out/c15_dws7_raw.csv 2>/dev/null
ncu -i /tmp/dws7.ncu-rep --page source --csv > gpurun_out/c15_dws7_source.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dw_small -s 1 -c 1 -f -o /tmp/dws14 python tools/prof_dw.py --C 480 --k 3 --s 1 --H 14 --frames 2048 --iters 2 > gpurun_out/c15_ncu_dws14.log 2>&1; echo "rc=$?"
ncu -i /tmp/dws14.ncu-rep --page raw --csv > gpurun_out/c15_dws14_raw.csv 2>/dev/null
ls -la gpurun_out | grep c15
