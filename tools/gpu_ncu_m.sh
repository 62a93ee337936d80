#!/bin/bash
# ncu full capture of the packed M-step kernels for three regimes: sparse list (c3 shape), dense list (W=12), scan (W=8)
mkdir -p gpurun_out
TAG=${1:-m}
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --nseq 300000"
i=0
for args in "--K 4" "--W 12 --K 2" "--W 8"; do
  i=$((i+1))
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_mstep_(list|scan)_w' -s 2 -c 2 -f -o gpurun_out/${TAG}_m$i $B $args > gpurun_out/${TAG}_m$i.log 2>&1
done
ls -la gpurun_out | tail -5
