#!/bin/bash
# quick A/B on the GPU box: bench variants + ncu launch list (kernel durations)
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-e2e --nseq 300000"
show() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$1', 'ms/step %.3f E %.3f M %.3f U %.3f' % (d['ms_per_step'], r['estep_ms'], r['mstep_accum_ms'], r['reduce_update_ms']))
"; }
$B | show default
BAMM_LIST_FRAC=0 $B | show nolist
BAMM_NO_REDUCED=1 $B | show noreduced
BAMM_NO_REDUCED=1 BAMM_LIST_FRAC=0 $B | show noreduced_nolist
BAMM_TABLE_BYTES=170000 BAMM_LIST_FRAC=0 $B | show table170k_nolist
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum --clock-control none -c 60 --csv --log-file gpurun_out/q_launches.csv $B > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/q_launches.csv')) if len(r)>10 and r[0].isdigit()]
agg={}
for r in rows:
    k=(r[4].split('(')[0][-40:], r[12])
    agg.setdefault(k,[]).append(float(r[14].replace(',','')))
for k,v in sorted(agg.items()):
    print(k, 'n=%d'%len(v), 'mean=%.4g'%(sum(v)/len(v)))
PY
