for v in ${@:-0 0x40000 0x80000 0 0x40000 0x80000}; do
  MANO_B200_VSKIN_VARIANT=$v python bench.py --no-extras --no-e2e --no-cpu-baseline --steps 6 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('variant', '$v', {k: round(x['ms'],3) for k,x in d['stages_ms'].items()}, 'step', round(d['ms_per_step'],3), 'fwd-only', round(d['forward_only']['stages_ms']['fused_fwd'],3))
"
done
