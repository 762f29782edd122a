#!/bin/bash
# session O: pipeline with two batches in flight (submit / collect): parity, faces/s; detector launch list at batch 64
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline_gpu.py tests/test_dropin_cpp.py -m gpu -q --timeout=200 > gpurun_out/r2o_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log; tail -15 gpurun_out/r2o_pytest.log | cut -c1-300
timeout 600 python tools/run_bench_pipeline.py 20 > gpurun_out/r2o_pipeline.json 2> gpurun_out/r2o_pipeline.err
echo "pipeline rc=$?"; tail -c 300 gpurun_out/r2o_pipeline.err; cut -c1-1500 gpurun_out/r2o_pipeline.json
FR_PIPE_SUB=32 timeout 600 python tools/run_bench_pipeline.py 20 > gpurun_out/r2o_pipeline_sub32.json 2>/dev/null
FR_PIPE_SUB=64 timeout 600 python tools/run_bench_pipeline.py 20 > gpurun_out/r2o_pipeline_sub64.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_detect_launches_b64.csv python tools/perf_nets.py --stages detect --det-batch 64 --reps 1 > /dev/null 2>&1
ls -la gpurun_out | tail -6
