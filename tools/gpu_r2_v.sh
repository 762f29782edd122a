#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for sc in f8 f16; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_search_unknown_launches_${sc}_1250k.csv python tools/prof_search_unknown.py 1250000 $sc > gpurun_out/r2v_$sc.log 2>&1
tail -2 gpurun_out/r2v_$sc.log
done
