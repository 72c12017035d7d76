#!/bin/bash
# A/B timing of builds of the library on ONE box: bash profiles/tools/lib_ab.sh <suffix> ... ("default" = the shipped build;
# other suffixes select 3dhandposeestimation_b200/libmano_b200_<suffix>.so through MANO_B200_LIB)
for rep in 1 2; do
for v in "$@"; do
  if [ "$v" = "default" ]; then unset MANO_B200_LIB; else export MANO_B200_LIB=$PWD/3dhandposeestimation_b200/libmano_b200_$v.so; fi
  python bench.py --no-extras --no-e2e --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('lib', '$v', {k: round(x['ms'],3) for k,x in d['stages_ms'].items()}, 'step', round(d['ms_per_step'],3), 'fwd-only', round(d['forward_only']['stages_ms']['fused_fwd'],3))
"
done
done
