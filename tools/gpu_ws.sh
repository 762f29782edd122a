#!/bin/bash
mkdir -p gpurun_out
FR_HALO=3 timeout 60 python -m pytest tests/test_embedder_gpu.py -m gpu -q --timeout 50 2>&1 | tail -3 | cut -c1-200 | tee gpurun_out/ws_pytest.log
for B in 256 32; do for H in 3 1; do
  echo -n "b=$B FR_HALO=$H "; FR_HALO=$H timeout 60 python tools/perf_nets.py --stages embed --emb-batch $B --reps 20 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms'],4), round(d['faces_per_s']))"
done; done 2>&1 | tee gpurun_out/ws_ab8.txt
