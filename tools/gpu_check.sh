#!/bin/bash
# Staged GPU check: each stage in its own process (a trapped kernel poisons the CUDA context) with a timeout.
# Usage (on the GPU box): bash tools/gpu_check.sh [stage ...]
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
stages=("$@")
[ ${#stages[@]} -eq 0 ] && stages=(simple tc path_simt path)
rc_all=0
for s in "${stages[@]}"; do
  case $s in
    simple)    cmd=(python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "preprocess or stem or dwconv or se_gate or simt") ;;
    tc)        cmd=(python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tcgen05 or head_gemm_pool") ;;
    path_simt) cmd=(env DFD_GEMM_IMPL=simt python -m pytest tests/test_gpu_path.py -m gpu -q) ;;
    path)      cmd=(python -m pytest tests/test_gpu_path.py -m gpu -q) ;;
    all)       cmd=(python -m pytest tests -m gpu -q -x) ;;
    bench)     cmd=(python bench.py --steps 5 --warmup 3) ;;
    smoke)     cmd=(python __graft_entry__.py smoke) ;;
    *) echo "unknown stage $s"; continue ;;
  esac
  echo "=== stage $s: ${cmd[*]}"
  timeout 900 "${cmd[@]}" > gpurun_out/check_$s.log 2>&1
  rc=$?
  echo "=== stage $s exit $rc"
  tail -n 25 gpurun_out/check_$s.log
  [ $rc -ne 0 ] && rc_all=1
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.csv 2>&1
exit $rc_all
