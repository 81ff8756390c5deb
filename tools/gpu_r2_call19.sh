#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "mbconv_fused" > gpurun_out/c19_t_f7.log 2>&1; echo "fused test rc=$?"; tail -15 gpurun_out/c19_t_f7.log
timeout 200 python tools/prof_fused7.py --k 5; timeout 200 python tools/prof_fused7.py --k 3
timeout 900 python -m pytest tests/test_gpu_path.py tests/test_gpu_parity_large.py -m gpu -q -x > gpurun_out/c19_t_path.log 2>&1; echo "path tests rc=$?"; tail -4 gpurun_out/c19_t_path.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c19_bench.json 2> gpurun_out/c19_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c19_bench.json')); print("default", d["ms_per_step"], d["steady"]["median_ms"], {k:(v["launches"], v["ms"]) for k,v in d["kernels"].items()})
PY
DFD_LIB_PATH=build/variants/libdfd_nofused7.so timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c19_bench_nof7.json 2> gpurun_out/c19_bench_nof7.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c19_bench_nof7.json')); print("nofused7", d["ms_per_step"], d["steady"]["median_ms"], {k:(v["launches"], v["ms"]) for k,v in d["kernels"].items()})
PY
