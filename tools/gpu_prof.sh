#!/bin/bash
# ncu: launch list + full capture of the fused scan kernel. Outputs under gpurun_out/.
mkdir -p gpurun_out
ROWS=${ROWS:-10000000}
echo "== bench" ; timeout 900 python bench.py --rows $ROWS --steps 20 --no-cpu-baseline --no-pipeline 2>&1 | tail -2 | tee gpurun_out/bench_prof_$ROWS.json
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$ROWS.csv \
    python bench.py --rows $ROWS --steps 2 --warmup 3 --ramp-s 0 --no-cpu-baseline --no-pipeline > gpurun_out/ncu_bench_$ROWS.log 2>&1
tail -3 gpurun_out/ncu_bench_$ROWS.log
echo "== ncu full"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:cosine_topk_coarse -s 2 -c 1 -f -o gpurun_out/coarse_$ROWS \
    python bench.py --rows $ROWS --steps 1 --warmup 3 --ramp-s 0 --no-cpu-baseline --no-pipeline > gpurun_out/ncu_full_$ROWS.log 2>&1
tail -3 gpurun_out/ncu_full_$ROWS.log
ls -la gpurun_out
