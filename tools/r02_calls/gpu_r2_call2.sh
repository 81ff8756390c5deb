#!/bin/bash
# Round 2, GPU call 2: full suite on the cleaned-up default path, every bench config, fused-kernel variants, GEMM role accounting,
# source-level ncu of the fused kernel, launch lists of the ViT and resnet50 engines.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name: $*"; timeout 600 "$@" > gpurun_out/c2_$name.log 2>&1; echo "=== $name exit $?"; tail -n ${TAILN:-6} gpurun_out/c2_$name.log; }
TAILN=12 run tests python -m pytest tests -m gpu -q -x
run smoke python __graft_entry__.py smoke
B="python bench.py --steps 20 --warmup 3"
run bench_c2 $B
run bench_c3 $B --config 3
run bench_c5 $B --config 5
run bench_ens $B --config ensemble
run bench_bf16 $B --precision bf16 --no-cpu-baseline
./build/mufu_probe > gpurun_out/c2_mufu.log 2>&1; cat gpurun_out/c2_mufu.log
for v in regA104 regA80 regB104 regB80 regC128 nob2 xr4; do
  TAILN=1 run var_$v env DFD_LIB_PATH=$PWD/build/variants/libdfd_$v.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline
  python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/c2_var_$v.log").read().strip().splitlines()[-1])
    print("VARIANT $v: ms/step", round(l["ms_per_step"], 3), "steady", l["steady"]["median_ms"], "fused ms", l["kernels"].get("expand_dwconv_fused", {}).get("ms"), "dw", l["kernels"]["dwconv_se_squeeze"]["ms"], "expand", l["kernels"]["gemm_expand"]["ms"])
except Exception as e:
    print("VARIANT $v: failed", e)
PY
done
for b in 1 2 3; do TAILN=1 run fused_b$b python tools/prof_fused.py --block $b --frames 2048; done
# GEMM role accounting (who waits for what): head + pool, the 1152->320 project conv, a 14x14 gated project conv, a 7x7 expand
TAILN=14 run gemm_head env DFD_GEMM_DBG=32 python tools/prof_gemm.py --K 320 --N 1280 --HW 49 --frames 2048 --gate 0 --res 0 --act 1 --pool 1 --iters 3
TAILN=14 run gemm_p320 env DFD_GEMM_DBG=32 python tools/prof_gemm.py --K 1152 --N 320 --HW 49 --frames 2048 --gate 1 --res 0 --act 0 --iters 3
TAILN=14 run gemm_p112 env DFD_GEMM_DBG=32 python tools/prof_gemm.py --K 672 --N 112 --HW 196 --frames 2048 --gate 1 --res 1 --act 0 --iters 3
TAILN=14 run gemm_e1152 env DFD_GEMM_DBG=32 python tools/prof_gemm.py --K 192 --N 1152 --HW 49 --frames 2048 --gate 0 --res 0 --act 1 --iters 3
# source-level profile of the fused kernel (block 2.1.0), 512 frames
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mbconv_fused -s 1 -c 1 -f -o /tmp/fused1 python tools/prof_fused.py --block 1 --frames 512 --iters 2 > gpurun_out/c2_ncu_fused1.log 2>&1
ncu -i /tmp/fused1.ncu-rep --page raw --csv > gpurun_out/c2_fused1_raw.csv 2>/dev/null
ncu -i /tmp/fused1.ncu-rep --page source --csv > gpurun_out/c2_fused1_source.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv_march -s 1 -c 1 -f -o /tmp/dw5 python tools/prof_dw.py --C 672 --k 5 --s 1 --H 14 --frames 512 --iters 2 > gpurun_out/c2_ncu_dw5.log 2>&1
ncu -i /tmp/dw5.ncu-rep --page raw --csv > gpurun_out/c2_dw5_raw.csv 2>/dev/null
ncu -i /tmp/dw5.ncu-rep --page source --csv > gpurun_out/c2_dw5_source.csv 2>/dev/null
# launch lists: ViT-B/16 (batch 128) and the resnet50 member (256 frames)
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_tensor.sum"
timeout 400 ncu --metrics $M --clock-control none -s 176 -c 88 --csv --log-file gpurun_out/c2_launches_vit.csv python tools/bench_vit.py --batch 128 --iters 1 > gpurun_out/c2_ncu_vit.log 2>&1
echo "vit launch list rc=$?"
timeout 400 ncu --metrics $M --clock-control none -k regex:"gemm_tc|rn_" -c 400 --csv --log-file gpurun_out/c2_launches_resnet.csv python tools/bench_resnet.py --videos 8 --frames 32 --iters 1 > gpurun_out/c2_ncu_resnet.log 2>&1
echo "resnet launch list rc=$?"
run parity python tools/parity_diag.py --out gpurun_out/parity_diag_r02.json
du -sh gpurun_out
