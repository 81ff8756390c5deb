#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c16_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c16_tests.log
timeout 300 python bench.py > gpurun_out/c16_bench.json 2> gpurun_out/c16_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c16_bench.json')); print("bench", d["value"], d["ms_per_step"], d["steady"]["median_ms"], d["e2e"]["value"], d["roofline"]["kernel"], d["roofline"]["frac"], {k:v["ms"] for k,v in d["kernels"].items()})
PY
timeout 400 python bench.py --config ensemble --no-cpu-baseline > gpurun_out/c16_bench_ens.json 2> gpurun_out/c16_bench_ens.err; cut -c1-300 gpurun_out/c16_bench_ens.json; grep -o '"roofline.*' gpurun_out/c16_bench_ens.json | cut -c1-400
timeout 300 python tools/bench_resnet.py --videos 8 --frames 32 --iters 3 > gpurun_out/c16_resnet_plain.log 2>&1; tail -1 gpurun_out/c16_resnet_plain.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/c16_launches_resnet.csv python tools/bench_resnet.py --videos 8 --frames 32 --iters 1 > gpurun_out/c16_ncu_resnet.log 2>&1; echo "resnet list rc=$?"
