#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/perf_nets.py --stages embed,e2e 2>&1 | tail -2
export FR_NO_GRAPHS=1
for B in 32 256; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_embed_b$B.csv \
    python tools/perf_nets.py --stages embed --emb-batch $B --reps 1 > /dev/null 2>&1
done
