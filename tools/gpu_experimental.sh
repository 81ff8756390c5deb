#!/bin/bash
# First GPU call of the next session: verify and time the paths that were written without GPU access (off by default).
# Each stage runs in its own process under a timeout (a trapped kernel poisons the CUDA context).
# Usage (on the GPU box): bash tools/gpu_experimental.sh
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name: $*"; timeout 600 "$@" > gpurun_out/exp_$name.log 2>&1; echo "=== $name exit $?"; tail -n 12 gpurun_out/exp_$name.log; }
# 1. implicit 3x3 convolution of the resnet50 member (gemm_tc.cu CONV variants, resnet.cu DFD_RESNET_IMPLICIT)
run conv_kernel  env DFD_EXPERIMENTAL=1 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k conv1x1_conv3x3
run conv_member  env DFD_EXPERIMENTAL=1 python -m pytest tests/test_resnet.py -m gpu -q -x -s -k implicit
run resnet_gather   python tools/bench_resnet.py --videos 8 --frames 32 --iters 5
run resnet_implicit env DFD_RESNET_IMPLICIT=1 python tools/bench_resnet.py --videos 8 --frames 32 --iters 5
# 2. programmatic dependent launch of the scoring step (engine: DFD_PDL=1), if present in this build
if grep -q DFD_PDL deepfake_video_detection_b200/csrc/api.cu 2>/dev/null; then
  run pdl_tests env DFD_PDL=1 python -m pytest tests/test_gpu_path.py -m gpu -q -x
  run pdl_off   python bench.py --steps 10 --warmup 3 --no-cpu-baseline
  run pdl_on    env DFD_PDL=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
fi
