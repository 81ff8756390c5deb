#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_vit.py -m gpu -q -s -k "attention_tcgen05" > gpurun_out/c7_t_attn.log 2>&1; echo "attn test rc=$?"; tail -25 gpurun_out/c7_t_attn.log
timeout 120 python tools/prof_vit_attn.py --images 512 2>&1 | tail -3
timeout 120 python tools/prof_vit_attn.py --images 128 2>&1 | tail -3
