#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_path.py -m gpu -q -x -k "stem or golden or determinism" > gpurun_out/c26_t.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/c26_t.log
for i in 1 2; do timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c26_bench.json 2> gpurun_out/c26_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c26_bench.json')); print("config2", d["ms_per_step"], d["steady"]["median_ms"], {k:v["ms"] for k,v in d["kernels"].items()})
PY
done
