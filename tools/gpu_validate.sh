#!/bin/bash
# One GPU round trip validating HEAD: parity tests, embedder under the three conv modes, the default bench. Outputs: gpurun_out/.
mkdir -p gpurun_out
T0=$(date +%s)
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi.txt 2>&1
echo "== embedder tests: halo (default)"; timeout 400 python -m pytest tests/test_embedder_gpu.py -m gpu -q --timeout 200 2>&1 | tail -8 | tee gpurun_out/emb_halo1.log
echo "== embedder tests: halo, base offset 0"; FR_HALO_BASEOFF=0 timeout 400 python -m pytest tests/test_embedder_gpu.py -m gpu -q --timeout 200 2>&1 | tail -8 | tee gpurun_out/emb_halo0.log
echo "== embedder tests: no halo"; FR_NO_HALO=1 timeout 400 python -m pytest tests/test_embedder_gpu.py -m gpu -q --timeout 200 2>&1 | tail -8 | tee gpurun_out/emb_nohalo.log
echo "t=$(( $(date +%s) - T0 ))s"
for B in 32 256; do
  echo "== perf embed b=$B halo";   timeout 200 python tools/perf_nets.py --stages embed --emb-batch $B 2>&1 | tail -1 | tee -a gpurun_out/perf_embed.json
  echo "== perf embed b=$B halo0";  FR_HALO_BASEOFF=0 timeout 200 python tools/perf_nets.py --stages embed --emb-batch $B 2>&1 | tail -1 | tee -a gpurun_out/perf_embed.json
  echo "== perf embed b=$B nohalo"; FR_NO_HALO=1 timeout 200 python tools/perf_nets.py --stages embed --emb-batch $B 2>&1 | tail -1 | tee -a gpurun_out/perf_embed.json
done
echo "t=$(( $(date +%s) - T0 ))s"
echo "== full gpu suite (FR_NO_HALO=${SUITE_NO_HALO:-1})"
FR_NO_HALO=1 timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -x --deselect tests/test_embedder_gpu.py 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "t=$(( $(date +%s) - T0 ))s"
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench 1.25M (8-GPU shard size on one GPU)"
FR_NO_HALO=1 timeout 600 python bench.py --rows 1250000 --steps 20 --no-cpu-baseline --no-pipeline 2>&1 | tail -1 | tee gpurun_out/bench_1250k.json
echo "== bench 10M default"
FR_NO_HALO=1 timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_10m.json
echo "t=$(( $(date +%s) - T0 ))s"
