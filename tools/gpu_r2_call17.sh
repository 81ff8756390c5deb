#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "head_gemm_pool" > gpurun_out/c17_t_head.log 2>&1; echo "head test rc=$?"; tail -12 gpurun_out/c17_t_head.log
timeout 900 python -m pytest tests/test_gpu_path.py tests/test_gpu_parity_large.py -m gpu -q -x > gpurun_out/c17_t_path.log 2>&1; echo "path tests rc=$?"; tail -4 gpurun_out/c17_t_path.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c17_bench.json 2> gpurun_out/c17_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c17_bench.json')); print("default", d["ms_per_step"], d["steady"]["median_ms"], {k:(v["ms"], v["TFLOPs"]) for k,v in d["kernels"].items()}); print(d["roofline"].get("fused_class"))
PY
