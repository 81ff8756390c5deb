#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c25_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c25_tests.log
S="--K 576 --N 64 --HW 3136 --frames 96 --gate 0 --res 0 --act 3"
python tools/prof_gemm.py $S | tail -1; DFD_GEMM_DBG=3 python tools/prof_gemm.py $S | tail -1
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c25_bench.json 2> gpurun_out/c25_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c25_bench.json')); print("config2", d["ms_per_step"], d["steady"]["median_ms"], {k:v["ms"] for k,v in d["kernels"].items()})
PY
timeout 300 python tools/bench_resnet.py --videos 8 --frames 32 --iters 5 | tail -1 | cut -c1-220
timeout 120 python tools/prof_vit_attn.py --images 512 | tail -1
timeout 300 python tools/prof_gemm_pair.py --images 512 --impls 3 2>&1 | tail -4
