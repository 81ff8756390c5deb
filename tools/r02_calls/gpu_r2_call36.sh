#!/bin/bash
# call 36: fused expand + depthwise kernel with the expanded ring kept in fp32 (no pack / unpack): parity, class time, carve-out A/B
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_path.py tests/test_gpu_parity_large.py -m gpu -q -x -k "fused or golden or determinism or graph or parity or large" > gpurun_out/c36_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c36_pytest.log
cp gpurun_out/parity_large_test.json gpurun_out/c36_parity_large_test.json 2>/dev/null
for v in default carve0 default; do
  if [ $v = default ]; then unset DFD_LIB_PATH; else export DFD_LIB_PATH=build/variants/libdfd_$v.so; fi
  timeout 120 python tools/time_classes.py --iters 3 2>&1 | tail -1
done
unset DFD_LIB_PATH
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c36_bench.json 2> gpurun_out/c36_bench.err; cut -c1-230 gpurun_out/c36_bench.json
