#!/bin/bash
# session U: dw3x3 with shared-memory weights (parity + launch lists); launch list of the search step on a 1.25 M-row shard (both scan copies)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_detector_gpu.py -m gpu -q --timeout=200 > gpurun_out/r2u_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2u_pytest.log; tail -4 gpurun_out/r2u_pytest.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_detect_launches_b64.csv python tools/perf_nets.py --stages detect --det-batch 64 --reps 1 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_detect_launches_b16.csv python tools/perf_nets.py --stages detect --reps 1 > /dev/null 2>&1
for sc in f8 f16; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_search_launches_${sc}_1250k.csv python bench.py --rows 1250000 --scan $sc --no-pipeline --no-cpu-baseline --no-ref-gpu --no-alt-scan --no-traffic --no-graph --no-unknown --steps 4 --warmup 3 --min-phase-s 0.02 --passes 1 > gpurun_out/r2u_search_$sc.json 2> gpurun_out/r2u_search_$sc.err
echo "search $sc rc=$?"
done
timeout 600 python tools/run_bench_pipeline.py 20 > gpurun_out/r2u_pipeline.json 2> gpurun_out/r2u_pipeline.err
echo "pipeline rc=$?"
