#!/bin/bash
# Round 2, GPU call 3: full suite (no -x), occupancy variants of the fused and 5x5 depthwise kernels, resnet50 member with the
# shared pool-head kernel.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name: $*"; timeout 900 "$@" > gpurun_out/c3_$name.log 2>&1; echo "=== $name exit $?"; tail -n ${TAILN:-6} gpurun_out/c3_$name.log; }
TAILN=15 run tests python -m pytest tests -m gpu -q
run smoke python __graft_entry__.py smoke
TAILN=1 run bench_c2 python bench.py --steps 20 --warmup 3 --no-cpu-baseline
for v in fcA fcB fcC fcD fcE dw5r128 dw5r144; do
  TAILN=1 run var_$v env DFD_LIB_PATH=$PWD/build/variants/libdfd_$v.so python bench.py --steps 20 --warmup 3 --no-cpu-baseline
done
python - <<'PY'
import json, glob
for f in ["gpurun_out/c3_bench_c2.log"] + sorted(glob.glob("gpurun_out/c3_var_*.log")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        k = l["kernels"]
        print(f.split("c3_")[1][:-4].ljust(14), "ms/step", round(l["ms_per_step"], 3), "steady", l["steady"]["median_ms"], "fused", k["expand_dwconv_fused"]["ms"], "dw", k["dwconv_se_squeeze"]["ms"],
              "expand", k["gemm_expand"]["ms"], "project", k["gemm_project"]["ms"], "stem", k["stem"]["ms"], "se", k["se_gate"]["ms"])
    except Exception as e:
        print(f, "failed", e)
PY
for b in 1 2 3; do for v in fcA fcB fcC fcD; do echo -n "$v "; DFD_LIB_PATH=$PWD/build/variants/libdfd_$v.so timeout 120 python tools/prof_fused.py --block $b --frames 2048 2>&1 | tail -1; done; done
for v in base dw5r128 dw5r144; do L=""; [ $v != base ] && L=$PWD/build/variants/libdfd_$v.so; for shape in "240 28" "480 14" "672 14" "1152 7"; do set -- $shape; echo -n "$v "; DFD_LIB_PATH=$L timeout 120 python tools/prof_dw.py --C $1 --k 5 --s 1 --H $2 --frames 2048 2>&1 | tail -1; done; done
TAILN=2 run resnet python tools/bench_resnet.py --videos 8 --frames 32 --iters 5
TAILN=1 run bench_ens python bench.py --steps 20 --warmup 3 --config ensemble
du -sh gpurun_out
