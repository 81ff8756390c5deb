#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "dwconv or se_gate" > gpurun_out/c13_t_k.log 2>&1; echo "kernel tests rc=$?"; tail -4 gpurun_out/c13_t_k.log
timeout 900 python -m pytest tests/test_gpu_path.py tests/test_gpu_parity_large.py -m gpu -q -x > gpurun_out/c13_t_path.log 2>&1; echo "path tests rc=$?"; tail -4 gpurun_out/c13_t_path.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c13_bench.json')); print("default", d["ms_per_step"], d["steady"]["median_ms"], {k:v["ms"] for k,v in d["kernels"].items()})
PY
for v in nodwsmall nosewide; do DFD_LIB_PATH=build/variants/libdfd_$v.so timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c13_bench_$v.json 2> gpurun_out/c13_bench_$v.err; V=$v python - <<'PY'
import json,os
v=os.environ["V"]; d=json.load(open(f'gpurun_out/c13_bench_{v}.json')); print(v, d["ms_per_step"], d["steady"]["median_ms"], {k:v["ms"] for k,v in d["kernels"].items()})
PY
done
