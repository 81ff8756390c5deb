#!/bin/bash
# Round 2, GPU call 4: restructured fused kernel (ldmatrix, bias in the accumulator, conflict-free ring pitch, phased expand, expand / depthwise overlap)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name: $*"; timeout 900 "$@" > gpurun_out/c4_$name.log 2>&1; echo "=== $name exit $?"; tail -n ${TAILN:-6} gpurun_out/c4_$name.log; }
TAILN=8 run tests python -m pytest tests/test_gpu_kernels.py tests/test_gpu_path.py tests/test_gpu_parity_large.py -m gpu -q
TAILN=1 run bench_c2 python bench.py --steps 20 --warmup 3 --no-cpu-baseline
for v in f2 f3c; do TAILN=1 run var_$v env DFD_LIB_PATH=$PWD/build/variants/libdfd_$v.so python bench.py --steps 20 --warmup 3 --no-cpu-baseline; done
python - <<'PY'
import json, glob
for f in ["gpurun_out/c4_bench_c2.log"] + sorted(glob.glob("gpurun_out/c4_var_*.log")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1]); k = l["kernels"]
        print(f.split("c4_")[1][:-4].ljust(14), "ms/step", round(l["ms_per_step"], 3), "steady", l["steady"]["median_ms"], "fused", k["expand_dwconv_fused"]["ms"], "dw", k["dwconv_se_squeeze"]["ms"],
              "expand", k["gemm_expand"]["ms"], "project", k["gemm_project"]["ms"], "stem", k["stem"]["ms"], "se", k["se_gate"]["ms"])
    except Exception as e:
        print(f, "failed", e)
PY
for b in 1 2 3; do for v in base f2 f3c; do L=""; [ $v != base ] && L=$PWD/build/variants/libdfd_$v.so; echo -n "$v "; DFD_LIB_PATH=$L timeout 120 python tools/prof_fused.py --block $b --frames 2048 2>&1 | tail -1; done; done
for shape in "240 28" "480 14" "672 14" "1152 7"; do set -- $shape; timeout 120 python tools/prof_dw.py --C $1 --k 5 --s 1 --H $2 --frames 2048 2>&1 | tail -1; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mbconv_fused -s 1 -c 1 -f -o /tmp/fused1 python tools/prof_fused.py --block 1 --frames 512 --iters 2 > gpurun_out/c4_ncu_fused1.log 2>&1
ncu -i /tmp/fused1.ncu-rep --page raw --csv > gpurun_out/c4_fused1_raw.csv 2>/dev/null
ncu -i /tmp/fused1.ncu-rep --page source --csv > gpurun_out/c4_fused1_source.csv 2>/dev/null
python tools/ncu_table.py gpurun_out/c4_fused1_raw.csv
du -sh gpurun_out
