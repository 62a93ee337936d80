#!/bin/bash
# Runs on the GPU box (via gpurun): shape + parity tests and a short c3 / c2 bench without the CPU baseline (a quick look after a kernel change).
set -u
mkdir -p gpurun_out
TAG=${1:-q}
timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_parity.py -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -25 | tee gpurun_out/${TAG}_pytest.txt
BAMM_DEBUG_LIST=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/${TAG}_c3.err | tee gpurun_out/${TAG}_c3.json
grep -m6 "pruned\|active list" gpurun_out/${TAG}_c3.err
timeout 300 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/${TAG}_c2.err | tee gpurun_out/${TAG}_c2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_ebound|k_eexact|k_emasked|k_estep_packed|k_mstep" -c 24 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/${TAG}_launches.csv')) if len(r)>5]
h=[i for i,r in enumerate(rows) if r[0]=='ID'][0]; H=rows[h]; ki=H.index('Kernel Name'); vi=H.index('Metric Value')
for r in rows[h+1:][-7:]: print(r[ki][:40], r[vi])
PY
