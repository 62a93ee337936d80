#!/bin/bash
# A/B of E-step build variants (build/variants/lib_*.so built with -DBAMM_E_THREADS / -DBAMM_E_UNROLL / -DBAMM_E_PIN)
mkdir -p gpurun_out
OUT=gpurun_out/${1:-variants}.txt; : > $OUT
for lib in bammmotif2_b200/libbamm_b200.so build/variants/lib_*.so; do
  for args in "--K 4" "--K 2" "--K 2 --W 12" "--K 2 --W 12 --L0 200"; do
  BAMM_LIB=$PWD/$lib python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --nseq ${NSEQ:-300000} $args 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$lib', '$args', 'ms/step %.3f E %.3f M %.3f U %.3f' % (d['ms_per_step'], r['estep_ms'], r['mstep_accum_ms'], r['reduce_update_ms']))
" | tee -a $OUT
  done
done
