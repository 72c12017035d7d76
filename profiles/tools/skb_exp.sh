#!/bin/bash
# removal experiments of skin_backward_kernel (MANO_B200_SKB_EXP bits, see skin.cu): bash profiles/tools/skb_exp.sh [bits ...]
for e in ${@:-0 1 2 3 4 8 16 24 7}; do
  MANO_B200_SKB_EXP=$e python bench.py --no-extras --no-e2e --no-cpu-baseline --steps 6 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('exp', $e, 'lbs_bwd', round(d['stages_ms']['lbs_bwd']['ms'],3), 'step', round(d['ms_per_step'],3))
"
done
