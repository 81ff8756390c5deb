#!/bin/bash
# Round 2, GPU call 1: verify / time every experimental path, large-sample parity diagnostic, launch lists.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name: $*"; timeout 420 "$@" > gpurun_out/c1_$name.log 2>&1; echo "=== $name exit $?"; tail -n 8 gpurun_out/c1_$name.log; }
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
run t_effnet env DFD_EXPERIMENTAL=1 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_path.py -m gpu -q -s -k "mbconv_fused or stem_dw_fused or se_gate_v2 or fused_expand"
run t_resnet env DFD_EXPERIMENTAL=1 python -m pytest tests/test_gpu_kernels.py tests/test_resnet.py -m gpu -q -s -k "conv1x1_conv3x3 or implicit"
run t_vit    env DFD_EXPERIMENTAL=1 python -m pytest tests/test_vit.py -m gpu -q -s -k "attention_v2 or epilogue_warps"
run bench_base   $B
run bench_fused1 env DFD_FUSE_EXPAND=1 $B
run bench_fused2 env DFD_FUSE_EXPAND=2 $B
run bench_fused3 env DFD_FUSE_EXPAND=3 $B
run bench_se2    env DFD_SE_V2=1 $B
run bench_all    env DFD_SE_V2=1 DFD_FUSE_EXPAND=3 $B
run bench_f1w    env DFD_FUSE_EXPAND=1 DFD_FUSE_CB=1 $B
run parity_base  python tools/parity_diag.py --out gpurun_out/parity_diag_base.json
run parity_fused env DFD_SE_V2=1 DFD_FUSE_EXPAND=3 python tools/parity_diag.py --precision fp16 --out gpurun_out/parity_diag_fused3.json
run path_fused env DFD_FUSE_EXPAND=3 DFD_SE_V2=1 python -m pytest tests/test_gpu_path.py -m gpu -q -x
run vit_base python tools/bench_vit.py --batch 512 --iters 5
run vit_att2 env DFD_VIT_ATTN_V2=1 python tools/bench_vit.py --batch 512 --iters 5
run vit_both env DFD_GEMM_F32_EPI16=1 DFD_VIT_ATTN_V2=1 python tools/bench_vit.py --batch 512 --iters 5
run rnn_epi16 env DFD_GEMM_F32_EPI16=1 python -m pytest tests/test_rnn.py -m gpu -q -x
run resnet_gather   python tools/bench_resnet.py --videos 8 --frames 32 --iters 5
run resnet_implicit env DFD_RESNET_IMPLICIT=1 python tools/bench_resnet.py --videos 8 --frames 32 --iters 5
run h2d python tools/probe_h2d.py
# launch lists (per-kernel device time + DRAM bytes) of one full-size step, default and everything-on
CMD="python tools/prof_step.py --videos 64 --frames 32 --iters 2"
for cfg in base all; do
  if [ $cfg = all ]; then export DFD_SE_V2=1 DFD_FUSE_EXPAND=3; fi
  timeout 200 $CMD > gpurun_out/c1_prof_$cfg.log 2>&1
  L=$(grep -o '[0-9]* launches' gpurun_out/c1_prof_$cfg.log | tail -1 | cut -d' ' -f1); L=${L:-71}
  timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s $L -c $L --csv --log-file gpurun_out/c1_launches_$cfg.csv $CMD > gpurun_out/c1_ncu_$cfg.log 2>&1
  echo "launch list $cfg rc=$? (L=$L)"
done
unset DFD_SE_V2 DFD_FUSE_EXPAND
du -sh gpurun_out
