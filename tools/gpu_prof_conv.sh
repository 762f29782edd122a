#!/bin/bash
mkdir -p gpurun_out
export FR_NO_GRAPHS=1
# stage-3 3x3 conv (grid 57x2) and stage-1 conv (grid 813): skip the first forward's launches
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 150 -c 40 -f -o gpurun_out/conv_embed \
    python tools/perf_nets.py --stages embed --reps 1 > gpurun_out/ncu_conv.log 2>&1
tail -2 gpurun_out/ncu_conv.log
ls -la gpurun_out/*.ncu-rep
