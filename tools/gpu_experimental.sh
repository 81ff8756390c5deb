#!/bin/bash
# First GPU call of the next session: verify and time the paths that were written without GPU access (all off by default).
# Each stage runs in its own process under a timeout (a trapped kernel poisons the CUDA context).
# Usage (on the GPU box): bash tools/gpu_experimental.sh
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name: $*"; timeout 600 "$@" > gpurun_out/exp_$name.log 2>&1; echo "=== $name exit $?"; tail -n 12 gpurun_out/exp_$name.log; }
# 0. is bench.py's e2e bound by the host->device copy or by the kernels?
run h2d python tools/probe_h2d.py
# 1. expand 1x1 fused into the marching depthwise kernel (mbconv_fused.cu, engine switch DFD_FUSE_EXPAND=1)
run fused_kernel env DFD_EXPERIMENTAL=1 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "mbconv_fused or stem_dw_fused"
run fused_path   env DFD_EXPERIMENTAL=1 python -m pytest tests/test_gpu_path.py -m gpu -q -x -s -k fused_expand
run fused_all    env DFD_FUSE_EXPAND=1 python -m pytest tests/test_gpu_path.py -m gpu -q -x
run bench_base   python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run bench_fused  env DFD_FUSE_EXPAND=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run bench_fused2 env DFD_FUSE_EXPAND=2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run bench_fused3 env DFD_FUSE_EXPAND=3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run fused_wide_t env DFD_FUSE_EXPAND=1 DFD_FUSE_CB=1 DFD_EXPERIMENTAL=1 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_path.py -m gpu -q -x -k "mbconv_fused or fused_expand"
run bench_fusedw env DFD_FUSE_EXPAND=1 DFD_FUSE_CB=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
# 2. implicit 3x3 convolution of the resnet50 member (gemm_tc.cu CONV variants, resnet.cu switch DFD_RESNET_IMPLICIT=1)
run conv_kernel  env DFD_EXPERIMENTAL=1 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k conv1x1_conv3x3
run conv_member  env DFD_EXPERIMENTAL=1 python -m pytest tests/test_resnet.py -m gpu -q -x -s -k implicit
run resnet_gather   python tools/bench_resnet.py --videos 8 --frames 32 --iters 5
run resnet_implicit env DFD_RESNET_IMPLICIT=1 python tools/bench_resnet.py --videos 8 --frames 32 --iters 5
# 3. ViT attention, second variant (vit.cu, DFD_VIT_ATTN_V2=1)
run vit_att2_test env DFD_EXPERIMENTAL=1 python -m pytest tests/test_vit.py -m gpu -q -x -s -k attention_v2
run vit_base      python tools/bench_vit.py --batch 512 --iters 5
run vit_att2      env DFD_VIT_ATTN_V2=1 python tools/bench_vit.py --batch 512 --iters 5
run vit_epi16_t   env DFD_EXPERIMENTAL=1 python -m pytest tests/test_vit.py tests/test_rnn.py -m gpu -q -x -k "epilogue_warps or rnn"
run vit_epi16     env DFD_GEMM_F32_EPI16=1 python tools/bench_vit.py --batch 512 --iters 5
run vit_both      env DFD_GEMM_F32_EPI16=1 DFD_VIT_ATTN_V2=1 python tools/bench_vit.py --batch 512 --iters 5
# 4. squeeze-excite gate, second variant (se.cu, DFD_SE_V2=1)
run se2_kernel env DFD_EXPERIMENTAL=1 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k se_gate_v2
run se2_path   env DFD_SE_V2=1 python -m pytest tests/test_gpu_path.py -m gpu -q -x
run bench_se2  env DFD_SE_V2=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run bench_all  env DFD_SE_V2=1 DFD_FUSE_EXPAND=3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
# 5. one ncu --set full capture of the new kernels (only after the runs above exited 0 without ncu)
if grep -q "passed" gpurun_out/exp_fused_path.log 2>/dev/null; then
  CMD="python tools/prof_step.py --videos 16 --frames 32 --iters 2"
  DFD_FUSE_EXPAND=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:mbconv_fused -c 6 -f -o /tmp/full_fused \
      env DFD_FUSE_EXPAND=1 $CMD > gpurun_out/exp_ncu_fused.log 2>&1
  ncu -i /tmp/full_fused.ncu-rep --page raw --csv > gpurun_out/exp_full_fused_raw.csv 2>/dev/null
  python tools/ncu_table.py gpurun_out/exp_full_fused_raw.csv > gpurun_out/exp_full_fused_table.txt 2>&1; cat gpurun_out/exp_full_fused_table.txt
  cp /tmp/full_fused.ncu-rep gpurun_out/ 2>/dev/null
fi
du -sh gpurun_out
