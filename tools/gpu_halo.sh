#!/bin/bash
mkdir -p gpurun_out
echo "== embedder tests FR_HALO=1"; FR_HALO=1 timeout 300 python -m pytest tests/test_embedder_gpu.py -m gpu -q --timeout 200 2>&1 | tail -3
for B in 32 256; do for H in 0 1 0 1; do
  echo -n "b=$B FR_HALO=$H "; FR_HALO=$H timeout 200 python tools/perf_nets.py --stages embed --emb-batch $B --reps 20 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms'],4), round(d['faces_per_s']))"
done; done 2>&1 | tee gpurun_out/halo_policy.txt
