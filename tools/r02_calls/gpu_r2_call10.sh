#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_pair -c 8 -f -o /tmp/pair python tools/prof_gemm_pair.py --images 512 --iters 1 --impls 3 > gpurun_out/c10_ncu_pair.log 2>&1; echo "ncu pair rc=$?"
ncu -i /tmp/pair.ncu-rep --page raw --csv > gpurun_out/c10_pair_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vit_attention -c 2 -f -o /tmp/att python tools/prof_vit_attn.py --images 512 > gpurun_out/c10_ncu_att.log 2>&1; echo "ncu att rc=$?"
ncu -i /tmp/att.ncu-rep --page raw --csv > gpurun_out/c10_att_raw.csv 2>/dev/null
ncu -i /tmp/att.ncu-rep --page source --csv > gpurun_out/c10_att_source.csv 2>/dev/null
ls -la gpurun_out | grep c10
