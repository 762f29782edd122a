#!/bin/bash
# ncu: full capture of the fused scan kernel on the e4m3 scan copy (10M and 1.25M rows). Outputs under gpurun_out/.
mkdir -p gpurun_out
for ROWS in 10000000 1250000; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cosine_topk_coarse -s 4 -c 1 -f -o gpurun_out/coarse_f8_$ROWS \
    python bench.py --rows $ROWS --scan f8 --steps 1 --warmup 3 --ramp-s 0 --no-cpu-baseline --no-pipeline --no-fp8 --no-graph > gpurun_out/ncu_full_f8_$ROWS.log 2>&1
tail -2 gpurun_out/ncu_full_f8_$ROWS.log
done
ls -la gpurun_out
