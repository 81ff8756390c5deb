#!/bin/bash
# call 37: depthwise march on one-strip maps without the padding columns' taps; full GPU suite on the stem / PDL changes
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c37_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c37_pytest.log
cp gpurun_out/parity_large_test.json gpurun_out/c37_parity_large_test.json 2>/dev/null
timeout 120 python tools/time_classes.py --iters 3 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dwconv_march -s 13 -c 13 --csv --log-file gpurun_out/c37_dw.csv python tools/prof_step.py --videos 64 --frames 32 > gpurun_out/c37_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'P'
import csv
rows=list(csv.reader(open('gpurun_out/c37_dw.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hi]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
for r in rows[hi+2:]:
    if len(r)>vi: print(r[ki][60:110], r[vi])
P
