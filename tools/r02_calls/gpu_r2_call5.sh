#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
./build/mufu_probe 2>&1 | tee gpurun_out/c5_mufu.log
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/c5_$name.log 2>&1; echo "=== $name exit $?"; }
run bench_c2 python bench.py --steps 20 --warmup 3 --no-cpu-baseline
for v in e80 e100; do run var_$v env DFD_LIB_PATH=$PWD/build/variants/libdfd_$v.so python bench.py --steps 20 --warmup 3 --no-cpu-baseline; done
python - <<'PY'
import json, glob
for f in ["gpurun_out/c5_bench_c2.log"] + sorted(glob.glob("gpurun_out/c5_var_*.log")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1]); k = l["kernels"]
        print(f.split("c5_")[1][:-4].ljust(14), "ms/step", round(l["ms_per_step"], 3), "steady", l["steady"]["median_ms"], "fused", k["expand_dwconv_fused"]["ms"], "dw", k["dwconv_se_squeeze"]["ms"],
              "expand", k["gemm_expand"]["ms"], "project", k["gemm_project"]["ms"], "stem", k["stem"]["ms"], "se", k["se_gate"]["ms"], "head", k["gemm_head_pool"]["ms"])
    except Exception as e:
        print(f, "failed", e)
PY
for v in base e80 e100; do L=""; [ $v != base ] && L=$PWD/build/variants/libdfd_$v.so; echo -n "$v "; DFD_LIB_PATH=$L timeout 120 python tools/prof_gemm.py --K 192 --N 1152 --HW 49 --frames 2048 --gate 0 --res 0 --act 1 --iters 5 2>&1 | tail -1; done
