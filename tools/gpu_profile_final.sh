#!/bin/bash
# Evidence run for profiles/: full tests, smoke, bench (+ reference arm), config-3 / config-5 timings, then the ncu launch
# list at the bench size and a --set full capture of every kernel class (CSV pages exported on the box).
# Usage: bash tools/gpu_profile_final.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$tag.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>/dev/null; cat gpurun_out/bench_ref_$tag.json
timeout 400 python bench.py --config 5 --no-cpu-baseline > gpurun_out/bench_vit_$tag.json 2> gpurun_out/bench_vit_$tag.err; cat gpurun_out/bench_vit_$tag.json
timeout 400 python bench.py --config 3 --no-cpu-baseline > gpurun_out/bench_rnn_$tag.json 2> gpurun_out/bench_rnn_$tag.err; cat gpurun_out/bench_rnn_$tag.json
timeout 400 python bench.py --config ensemble --no-cpu-baseline > gpurun_out/bench_ens_$tag.json 2> gpurun_out/bench_ens_$tag.err; cat gpurun_out/bench_ens_$tag.json
CMD="python tools/prof_step.py --videos 64 --frames 32 --iters 2"
timeout 300 $CMD > gpurun_out/prof_plain_$tag.log 2>&1
L=$(grep -o '[0-9]* launches' gpurun_out/prof_plain_$tag.log | tail -1 | cut -d' ' -f1); L=${L:-71}     # launches per forward pass (71 by default, fewer with DFD_FUSE_EXPAND)
echo "launches per pass: $L"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s $L -c $L --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_list_$tag.log 2>&1
echo "launch list rc=$?"
CMDV="python tools/bench_vit.py --batch 512 --iters 1"
timeout 300 $CMDV > /dev/null 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_tensor.sum --clock-control none -s 176 -c 88 --csv --log-file gpurun_out/launches_vit_$tag.csv $CMDV > gpurun_out/ncu_list_vit_$tag.log 2>&1
echo "vit launch list rc=$?"
timeout 300 python tools/prof_gemm_pair.py --images 512 > gpurun_out/prof_gemm_pair_$tag.log 2>&1; cat gpurun_out/prof_gemm_pair_$tag.log
timeout 120 python tools/prof_vit_attn.py --images 512 > gpurun_out/prof_vit_attn_$tag.log 2>&1; cat gpurun_out/prof_vit_attn_$tag.log
timeout 300 python tools/bench_resnet.py --videos 8 --frames 32 --iters 3 > gpurun_out/bench_resnet_$tag.json 2>&1; tail -1 gpurun_out/bench_resnet_$tag.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_resnet_$tag.csv python tools/bench_resnet.py --videos 8 --frames 32 --iters 1 > gpurun_out/ncu_list_resnet_$tag.log 2>&1
echo "resnet launch list rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:"gemm_pair|vit_attention|vit_layernorm" -s 30 -c 8 -f -o /tmp/fullvit_$tag $CMDV > gpurun_out/ncu_fullvit_$tag.log 2>&1
ncu -i /tmp/fullvit_$tag.ncu-rep --page raw --csv > gpurun_out/fullvit_${tag}_raw.csv 2>/dev/null
echo "vit full rc=$?"
CMD2="python tools/prof_step.py --videos 16 --frames 32 --iters 2"
timeout 300 $CMD2 > /dev/null 2>&1 && timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"dwconv|mbconv_fused|gemm_tc|stem|se_kernel|se_wide|pool_head|head_pool|scale_weights" -s $L -c $L -f -o /tmp/full_$tag $CMD2 > gpurun_out/ncu_full_$tag.log 2>&1
echo "full rc=$?"
ncu -i /tmp/full_$tag.ncu-rep --page raw --csv > gpurun_out/full_${tag}_raw.csv 2>/dev/null
du -sh gpurun_out
