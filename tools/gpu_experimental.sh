#!/bin/bash
# First GPU call of the next session: verify and time the paths that were written without GPU access (all off by default).
# Each stage runs in its own process under a timeout (a trapped kernel poisons the CUDA context).
# Usage (on the GPU box): bash tools/gpu_experimental.sh [quick|full]      quick (default): ~12 stages, about 12 minutes of box time
set -u
mode=${1:-quick}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name: $*"; timeout 600 "$@" > gpurun_out/exp_$name.log 2>&1; echo "=== $name exit $?"; tail -n 12 gpurun_out/exp_$name.log; }
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"

# ---- parity of every experimental kernel: one process per kernel family
run t_effnet env DFD_EXPERIMENTAL=1 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_path.py -m gpu -q -s -k "mbconv_fused or stem_dw_fused or se_gate_v2 or fused_expand"
run t_resnet env DFD_EXPERIMENTAL=1 python -m pytest tests/test_gpu_kernels.py tests/test_resnet.py -m gpu -q -s -k "conv1x1_conv3x3 or implicit"
run t_vit    env DFD_EXPERIMENTAL=1 python -m pytest tests/test_vit.py -m gpu -q -s -k "attention_v2 or epilogue_warps"
# ---- headline bench: default, fused levels, everything on
run bench_base   $B
run bench_fused1 env DFD_FUSE_EXPAND=1 $B
run bench_fused3 env DFD_FUSE_EXPAND=3 $B
run bench_all    env DFD_SE_V2=1 DFD_FUSE_EXPAND=3 $B
# ---- ViT (config 5) and the resnet50 member
run vit_base python tools/bench_vit.py --batch 512 --iters 5
run vit_both env DFD_GEMM_F32_EPI16=1 DFD_VIT_ATTN_V2=1 python tools/bench_vit.py --batch 512 --iters 5
run resnet_gather   python tools/bench_resnet.py --videos 8 --frames 32 --iters 5
run resnet_implicit env DFD_RESNET_IMPLICIT=1 python tools/bench_resnet.py --videos 8 --frames 32 --iters 5
# ---- is bench.py's e2e bound by the host->device copy or by the kernels?
run h2d python tools/probe_h2d.py

if [ "$mode" = full ]; then
  # whole GPU path suite with the switches on
  run path_fused env DFD_FUSE_EXPAND=3 DFD_SE_V2=1 python -m pytest tests/test_gpu_path.py -m gpu -q -x
  run bench_fused2 env DFD_FUSE_EXPAND=2 $B
  run bench_se2    env DFD_SE_V2=1 $B
  # wider channel blocks of the fused kernel
  run fused_wide_t env DFD_FUSE_EXPAND=1 DFD_FUSE_CB=1 DFD_EXPERIMENTAL=1 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_path.py -m gpu -q -x -k "mbconv_fused or fused_expand"
  run bench_fusedw env DFD_FUSE_EXPAND=1 DFD_FUSE_CB=1 $B
  run vit_att2  env DFD_VIT_ATTN_V2=1 python tools/bench_vit.py --batch 512 --iters 5
  run vit_epi16 env DFD_GEMM_F32_EPI16=1 python tools/bench_vit.py --batch 512 --iters 5
  run rnn_epi16 env DFD_GEMM_F32_EPI16=1 python -m pytest tests/test_rnn.py -m gpu -q -x
  # one ncu --set full capture of the fused kernel (only after its tests passed without ncu)
  if grep -q "passed" gpurun_out/exp_t_effnet.log 2>/dev/null && ! grep -q "failed" gpurun_out/exp_t_effnet.log; then
    CMD="python tools/prof_step.py --videos 16 --frames 32 --iters 2"
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:mbconv_fused -c 8 -f -o /tmp/full_fused \
        env DFD_FUSE_EXPAND=3 $CMD > gpurun_out/exp_ncu_fused.log 2>&1
    ncu -i /tmp/full_fused.ncu-rep --page raw --csv > gpurun_out/exp_full_fused_raw.csv 2>/dev/null
    python tools/ncu_table.py gpurun_out/exp_full_fused_raw.csv > gpurun_out/exp_full_fused_table.txt 2>&1; cat gpurun_out/exp_full_fused_table.txt
    cp /tmp/full_fused.ncu-rep gpurun_out/ 2>/dev/null
  fi
fi
du -sh gpurun_out
