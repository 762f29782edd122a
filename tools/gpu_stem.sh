#!/bin/bash
mkdir -p gpurun_out
for B in 256 32; do for P in 0 1 0 1; do
  echo -n "b=$B FR_STEM_PAIR=$P "; FR_STEM_PAIR=$P timeout 200 python tools/perf_nets.py --stages embed --emb-batch $B --reps 20 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms'],4), round(d['faces_per_s']))"
done; done 2>&1 | tee gpurun_out/stem_pair.txt
