#!/bin/bash
# call 34: stem row kernel with two output rows per tile; ring depths; skeleton timings
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_path.py -m gpu -q -x -k "stem or golden or determinism or graph or u8" > gpurun_out/c34_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c34_pytest.log
for v in default stemS8R8 stemS6R12 stemS8R16 stemdbg15 stemdbg7 default; do
  if [ $v = default ]; then unset DFD_LIB_PATH; else export DFD_LIB_PATH=build/variants/libdfd_$v.so; fi
  timeout 120 python tools/time_classes.py --only stem --iters 3 2>&1 | tail -1
done
