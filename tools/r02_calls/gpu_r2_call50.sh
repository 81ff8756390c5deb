#!/bin/bash
# call 50: fused expand + depthwise kernel with TWO rows per CTA barrier (4 expanded row slots): parity and class time
set -u
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_path.py -m gpu -q -x -k "fused or golden or determinism or full_c2" 2>&1 | tail -2
timeout 120 python tools/time_classes.py --iters 5 2>&1 | tail -1
