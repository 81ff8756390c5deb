#!/bin/bash
# call 31: stem row kernel with merged hi/lo accumulators + 32-bit tile counters; PDL trigger placement (explicit at the top of
# every kernel vs implicit at exit) in a stream and under graph replay
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_path.py -m gpu -q -x -k "stem or golden or determinism or graph or u8" > gpurun_out/c31_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c31_pytest.log
for v in default notrig nopdl default; do
  if [ $v = default ]; then unset DFD_LIB_PATH; else export DFD_LIB_PATH=build/variants/libdfd_$v.so; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c31_bench_$v.json 2> gpurun_out/c31_bench_$v.err
  python - gpurun_out/c31_bench_$v.json $v <<'P'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[2], d["ms_per_step"], d["value"], "steady", d["steady"]["median_ms"], "e2e", d["e2e"]["value"], "stem", d["kernels"]["stem"], "clk", d["clocks"]["sm_mhz"])
P
done
unset DFD_LIB_PATH
echo "--- sweep notrig"; DFD_LIB_PATH=build/variants/libdfd_notrig.so timeout 300 python tools/sweep_batch.py 2>&1 | head -11
echo "--- sweep default"; timeout 300 python tools/sweep_batch.py 2>&1 | head -11
