#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"\(bool\)0, \(int\)2>" -s 13 -c 4 -f -o /tmp/conv2 python tools/bench_resnet.py --videos 8 --frames 32 --iters 1 > gpurun_out/c21_ncu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/c21_ncu.log
ncu -i /tmp/conv2.ncu-rep --page raw --csv > gpurun_out/c21_conv2_raw.csv 2>/dev/null
ncu -i /tmp/conv2.ncu-rep --page source --csv > gpurun_out/c21_conv2_source.csv 2>/dev/null
ls -la gpurun_out | grep c21
