#!/bin/bash
# Runs on the GPU box (via gpurun): bench lines (c3 with the CPU baseline, c2), reference arm, ncu launch list and one full capture of
# every hot kernel of one c3 iteration. Everything that should come back goes to gpurun_out/ with the tag as prefix.
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== bench c3"; timeout 1500 python bench.py --steps 10 --warmup 3 2>gpurun_out/${TAG}_bench_c3.err | tee gpurun_out/${TAG}_bench_c3.json | cut -c1-400
echo "== bench c2"; timeout 600 python bench.py --workload c2 --steps 20 --warmup 3 2>gpurun_out/${TAG}_bench_c2.err | tee gpurun_out/${TAG}_bench_c2.json | cut -c1-300
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_ref.json | cut -c1-300
echo "== ncu launch list"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_c3.csv \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_launches_bench.log 2>&1
echo "== ncu full (one iteration, c3)"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_emasked|k_ebound|k_eexact|k_mstep_list_w|k_reduce_parts|k_update_model' -s 12 -c 6 -f -o gpurun_out/${TAG}_prof \
    python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_prof_bench.log 2>&1
ls -la gpurun_out | tail -12
