#!/bin/bash
# accuracy + speed of library variants: bash tools/variant_eval.sh lib1.so lib2.so ...
for lib in "" "$@"; do
  echo "== ${lib:-default}"
  export DFD_LIB_PATH=${lib:+$PWD/$lib}
  [ -z "$lib" ] && unset DFD_LIB_PATH
  python tests/diag_numerics.py 99 2>&1 | python -c "import sys,json; d=json.load(sys.stdin); print('fp16 seed99', {k: round(v,5) if isinstance(v,float) else v for k,v in d['fp16'].items() if k in ('feat_rel','dlogit_max','dlogit_med','flips')})"
  python tests/diag_numerics.py 7 2>&1 | python -c "import sys,json; d=json.load(sys.stdin); print('fp16 seed7 ', {k: round(v,5) if isinstance(v,float) else v for k,v in d['fp16'].items() if k in ('feat_rel','dlogit_max','dlogit_med','flips')})"
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', round(d['value']), {k: v['ms'] for k,v in d['kernels'].items()})"
done
