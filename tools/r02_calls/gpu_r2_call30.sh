#!/bin/bash
# call 30: programmatic dependent launch on every kernel of the EfficientNet step — parity, A/B timing against the DFD_PDL=0 build,
# small-batch sweep (eager and graph replay), per-kernel times of an 8-frame step
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/c30_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c30_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c30_bench_pdl.json 2> gpurun_out/c30_bench_pdl.err; cut -c1-260 gpurun_out/c30_bench_pdl.json; grep -o '"kernels".*' gpurun_out/c30_bench_pdl.json | cut -c1-900
DFD_LIB_PATH=build/variants/libdfd_nopdl.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c30_bench_nopdl.json 2> gpurun_out/c30_bench_nopdl.err; cut -c1-260 gpurun_out/c30_bench_nopdl.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c30_bench_pdl2.json 2> gpurun_out/c30_bench_pdl2.err; cut -c1-260 gpurun_out/c30_bench_pdl2.json
echo "--- sweep PDL"; timeout 300 python tools/sweep_batch.py 2>&1 | head -11
echo "--- sweep no PDL"; DFD_LIB_PATH=build/variants/libdfd_nopdl.so timeout 300 python tools/sweep_batch.py 2>&1 | head -11
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 68 -c 68 --csv --log-file gpurun_out/c30_launches_f8.csv python tools/prof_step.py --videos 1 --frames 8 > gpurun_out/c30_ncu_f8.log 2>&1; echo "ncu rc=$?"
