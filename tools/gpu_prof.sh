#!/bin/bash
# plain run -> ncu launch list -> ncu --set full of selected kernels; CSV pages are exported ON the box so that
# gpurun_out stays small.  Usage: bash tools/gpu_prof.sh <tag> [kernel-regex] [count]
set -u
tag=${1:-r1}; regex=${2:-}; count=${3:-66}; skip=${4:-66}
mkdir -p gpurun_out
CMD="python tools/prof_step.py --videos 16 --frames 32 --iters 2"
$CMD > gpurun_out/prof_plain_$tag.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/prof_plain_$tag.log; exit 1; }
cat gpurun_out/prof_plain_$tag.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 66 -c 66 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_list_$tag.log 2>&1
echo "launch list rc=$?"
if [ -n "$regex" ]; then K="-k regex:$regex"; else K=""; fi
ncu --set full --clock-control none --import-source on $K -s $skip -c $count -f -o /tmp/prof_$tag $CMD > gpurun_out/ncu_full_$tag.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/ncu_full_$tag.log
ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
ls -la /tmp/prof_$tag.ncu-rep
sz=$(stat -c %s /tmp/prof_$tag.ncu-rep); if [ "$sz" -lt 40000000 ]; then cp /tmp/prof_$tag.ncu-rep gpurun_out/; fi
du -sh gpurun_out
