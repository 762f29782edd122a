#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/perf_nets.py 2>&1 | tail -5 | tee gpurun_out/perf_nets.json
echo "== ncu launches (embed)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_embed.csv \
    python tools/perf_nets.py --stages embed --reps 1 > gpurun_out/ncu_embed.log 2>&1
tail -2 gpurun_out/ncu_embed.log
echo "== ncu launches (detect)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_detect.csv \
    python tools/perf_nets.py --stages detect --reps 1 > gpurun_out/ncu_detect.log 2>&1
tail -2 gpurun_out/ncu_detect.log
