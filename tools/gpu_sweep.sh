#!/bin/bash
# order / width sweep of the EM iteration at the c3 shape (300k sequences per run): ms per step and the E/M split
mkdir -p gpurun_out
OUT=gpurun_out/${1:-sweep}.txt; : > $OUT
run() {
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --nseq ${NSEQ:-300000} "$@" 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; c=d['config']
        print('$*', 'ms/step %.3f E %.3f M %.3f U %.3f  pos.iter/s %.3e  frac12B %.3f' % (d['ms_per_step'], r['estep_ms'], r['mstep_accum_ms'], r['reduce_update_ms'], c['positions_iter_per_s'], r['whole_iteration_frac_of_12B_roofline']))
" | tee -a $OUT
}
for K in 0 1 2 3 4 5 6; do run --K $K; done
for W in 8 12 16 24 31; do run --W $W; done
run --K 5 --W 12
run --K 5 --W 16
run --K 3 --W 12
run --K 2 --W 12
run --L0 100
run --L0 200
