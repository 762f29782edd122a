#!/bin/bash
# session T: register-weight dw3x3 + occupancy grids: parity, pipeline faces/s, detector launch lists
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_detector_gpu.py tests/test_pipeline_gpu.py tests/test_dropin_cpp.py -m gpu -q --timeout=200 > gpurun_out/r2t_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log; tail -6 gpurun_out/r2t_pytest.log | cut -c1-300
timeout 600 python tools/run_bench_pipeline.py 20 > gpurun_out/r2t_pipeline.json 2> gpurun_out/r2t_pipeline.err
echo "pipeline rc=$?"; tail -c 300 gpurun_out/r2t_pipeline.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_detect_launches_b16.csv python tools/perf_nets.py --stages detect --reps 1 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_detect_launches_b64.csv python tools/perf_nets.py --stages detect --det-batch 64 --reps 1 > /dev/null 2>&1
