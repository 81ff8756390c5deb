#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S="--K 576 --N 64 --HW 3136 --frames 96 --gate 0 --res 0 --act 3"
echo "== plain"; python tools/prof_gemm.py $S | tail -1
echo "== skip loads (dbg 1)"; DFD_GEMM_DBG=1 python tools/prof_gemm.py $S | tail -1
echo "== skip MMAs (dbg 2)"; DFD_GEMM_DBG=2 python tools/prof_gemm.py $S | tail -1
echo "== skip loads + MMAs (dbg 3)"; DFD_GEMM_DBG=3 python tools/prof_gemm.py $S | tail -1
echo "== skip loads + MMAs + stores (dbg 7)"; DFD_GEMM_DBG=7 python tools/prof_gemm.py $S | tail -1
echo "== role waits (dbg 32)"; DFD_GEMM_DBG=32 python tools/prof_gemm.py $S --iters 1 | grep -E "mma|loader|epilogue" | sort | uniq -c | sort -rn | head -8
S2="--K 576 --N 256 --HW 3136 --frames 96 --gate 0 --res 0 --act 3"
echo "== N=256 plain"; python tools/prof_gemm.py $S2 | tail -1
echo "== N=256 skip loads"; DFD_GEMM_DBG=1 python tools/prof_gemm.py $S2 | tail -1
