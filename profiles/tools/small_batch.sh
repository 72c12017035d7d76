#!/bin/bash
# small-batch latency: BASELINE config 2 (B = 4096) and the reference's own training batch sizes
for B in 64 480 4096 16384; do
  python bench.py --hands $B --rotate 16 --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print($B, 'ms/step', round(d['ms_per_step'],4), {k: round(v['ms']*1e3,1) for k,v in d['stages_ms'].items()}, 'e2e', d['e2e']['value'])"
done
