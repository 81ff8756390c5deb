#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "dwconv or se_gate" > gpurun_out/c14_t_k.log 2>&1; echo "kernel tests rc=$?"; tail -4 gpurun_out/c14_t_k.log
for lib in "" build/variants/libdfd_nodwsmall.so; do
  echo "== dw lib: ${lib:-default (small-map kernel)}"
  for cfg in "480 3 14" "480 5 14" "672 5 14" "1152 5 7" "1152 3 7"; do set -- $cfg; DFD_LIB_PATH=$lib timeout 120 python tools/prof_dw.py --C $1 --k $2 --s 1 --H $3 --frames 2048 | tail -1; done
done
for lib in "" build/variants/libdfd_nosewide.so; do
  echo "== se lib: ${lib:-default (wide kernel)}"
  for cfg in "480 20 2" "672 28 2" "672 28 1" "1152 48 1"; do set -- $cfg; DFD_LIB_PATH=$lib timeout 120 python tools/prof_se.py --C $1 --rd $2 --nparts $3 | tail -1; done
done
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c14_bench.json 2> gpurun_out/c14_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c14_bench.json')); print("default", d["ms_per_step"], d["steady"]["median_ms"], {k:v["ms"] for k,v in d["kernels"].items()})
PY
