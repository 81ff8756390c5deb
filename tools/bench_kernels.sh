#!/bin/bash
# bench.py under env variants; prints step time and per-kernel-class breakdown.  Usage: bash tools/bench_kernels.sh "A=1" "B=2" ...
for env in "$@"; do
  env $env timeout 200 python bench.py --steps 5 --warmup 3 > /tmp/b.json 2>/dev/null
  python - "$env" <<'PY'
import json,sys
d=json.loads(open('/tmp/b.json').read().strip().split('\n')[-1])
print(sys.argv[1], round(d['ms_per_step'],3), {k:round(v['ms'],3) for k,v in d['kernels'].items()})
PY
done
