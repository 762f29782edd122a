#!/bin/bash
# One GPU round trip: smoke, GPU parity tests, a short bench. Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
echo "== pytest" ; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== bench 1M" ; timeout 600 python bench.py --rows 1000000 --steps 10 --no-cpu-baseline --no-pipeline 2>&1 | tail -3 | tee gpurun_out/bench_1m.json
echo "== bench 10M" ; timeout 900 python bench.py --steps 10 2>&1 | tail -3 | tee gpurun_out/bench_10m.json
