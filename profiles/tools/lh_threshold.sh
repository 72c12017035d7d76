#!/bin/bash
# historical: the MB_TUNE_LH_MIN override this sweep used was removed once LH_MIN_HANDS = 8192 (common.cuh) was fixed;
# to repeat it, change the constant and rebuild.  Results: profiles/r1/small_batch.md.
for T in 4096 1000000; do for B in 4096 8192 16384 32768 65536; do
  MB_TUNE_LH_MIN=$T python bench.py --hands $B --rotate 8 --steps 100 --warmup 10 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('lh_min', $T, 'B', $B, 'ms/step', round(d['ms_per_step'],4), {k: round(v['ms']*1e3,1) for k,v in d['stages_ms'].items()})"
done; done
