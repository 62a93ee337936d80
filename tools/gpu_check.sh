#!/bin/bash
# Runs on the GPU box (via gpurun): GPU parity tests, smoke, bench lines, microbench. No ncu (see gpu_round2.sh).
set -u
mkdir -p gpurun_out
TAG=${1:-chk}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/${TAG}_smoke.txt
echo "== bench c3"; timeout 1500 python bench.py --steps 10 --warmup 3 2>gpurun_out/${TAG}_bench_c3.err | tee gpurun_out/${TAG}_bench_c3.json
tail -5 gpurun_out/${TAG}_bench_c3.err
echo "== bench c2"; timeout 600 python bench.py --workload c2 --steps 20 --warmup 3 2>gpurun_out/${TAG}_bench_c2.err | tee gpurun_out/${TAG}_bench_c2.json
[ -x tools/microbench ] && { echo "== microbench"; timeout 300 tools/microbench | tee gpurun_out/${TAG}_microbench.txt; }
