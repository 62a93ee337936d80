#!/bin/bash
# config 4 (FDR data path): bench line + ncu launch list of a 100k-positive run (kernel durations of sampler / packer / scorer)
mkdir -p gpurun_out
TAG=${1:-c4}
python bench.py --workload c4 --steps 5 --warmup 3 2>gpurun_out/${TAG}_bench_c4.err | tee gpurun_out/${TAG}_bench_c4.json | cut -c1-600
python bench.py --workload c4 --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_c4_ref.json | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_c4.csv \
    python bench.py --workload c4 --nseq 100000 --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv,sys,os
tag=sys.argv[1] if len(sys.argv)>1 else os.environ.get("TAG","c4")
PY
python -c "
import csv
rows=[r for r in csv.reader(open('gpurun_out/${TAG}_launches_c4.csv')) if len(r)>10 and r[0].isdigit()]
agg={}
for r in rows:
    k=r[4].split('(')[0][-48:]
    agg.setdefault(k,[]).append(float(r[-1].replace(',','')))
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
    print('%-50s n=%3d total %.3f ms' % (k, len(v), sum(v)/1e6 if max(v)>1e4 else sum(v)))
"
