#!/bin/bash
# Runs on the GPU box (via gpurun): the EM parity tests and a short c3 / c2 bench without the CPU baseline (a quick look after a kernel change).
set -u
mkdir -p gpurun_out
TAG=${1:-q}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "first_iteration or optimize or subset_fold or kernel_paths or properties or large_tables or stepwise" 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/${TAG}_c3.err | tee gpurun_out/${TAG}_c3.json
timeout 300 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/${TAG}_c2.err | tee gpurun_out/${TAG}_c2.json
