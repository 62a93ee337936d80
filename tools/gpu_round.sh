#!/bin/bash
# Runs on the GPU box (via gpurun): GPU parity tests, smoke, bench, ncu launch list and one full capture.
# Everything that should come back goes to gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/${TAG}_smoke.txt
echo "== bench c3"; timeout 1500 python bench.py --steps 10 --warmup 3 2>gpurun_out/${TAG}_bench_c3.err | tee gpurun_out/${TAG}_bench_c3.json
tail -5 gpurun_out/${TAG}_bench_c3.err
echo "== bench c2"; timeout 600 python bench.py --workload c2 --steps 20 --warmup 3 2>gpurun_out/${TAG}_bench_c2.err | tee gpurun_out/${TAG}_bench_c2.json
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -3 | tee gpurun_out/${TAG}_bench_ref.json
echo "== ncu launch list"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_launches_bench.log 2>&1
echo "== ncu full (E-step and M-step, c3)"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_estep_packed|k_mstep_list_w' -s 2 -c 2 -f -o gpurun_out/${TAG}_prof \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_prof_bench.log 2>&1
ls -la gpurun_out
