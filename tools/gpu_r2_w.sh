#!/bin/bash
# session W: fp16 pre-filter in the e4m3 re-rank: search parity suite, unknown-query launch list, 1-GPU bench at 1.25 M rows
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_search_gpu.py tests/test_search_reference_lib.py -m gpu -q --timeout=600 > gpurun_out/r2w_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2w_pytest.log; tail -6 gpurun_out/r2w_pytest.log | cut -c1-300
for sc in f8; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_search_unknown_launches_${sc}_1250k.csv python tools/prof_search_unknown.py 1250000 $sc > gpurun_out/r2w_$sc.log 2>&1
tail -2 gpurun_out/r2w_$sc.log
done
timeout 600 python bench.py --scan f8 --no-pipeline --no-cpu-baseline --no-ref-gpu --no-traffic > gpurun_out/r2w_bench_f8.json 2> gpurun_out/r2w_bench_f8.err
echo "bench rc=$?"; tail -c 200 gpurun_out/r2w_bench_f8.err
