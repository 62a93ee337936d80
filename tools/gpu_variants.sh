#!/bin/bash
# A/B of E-step build variants (build/variants/lib_*.so built with -DBAMM_E_THREADS / -DBAMM_E_PIN)
B="python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-e2e --nseq ${NSEQ:-300000}"
for lib in build/variants/lib_*.so; do
  BAMM_LIB=$PWD/$lib $B | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$lib', 'ms/step %.3f E %.3f M %.3f U %.3f' % (d['ms_per_step'], r['estep_ms'], r['mstep_accum_ms'], r['reduce_update_ms']))
"
done
