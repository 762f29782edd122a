#!/bin/bash
# session Q (PDL on the embedder chain included): split-precision mma.sync weights, 32-bit detector indexing, unified tile policy: parity, perf, pipeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_detector_gpu.py tests/test_embedder_gpu.py tests/test_pipeline_gpu.py tests/test_dropin_cpp.py -m gpu -q --timeout=200 > gpurun_out/r2q_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log; tail -12 gpurun_out/r2q_pytest.log | cut -c1-300
run() { echo "== $1 $2" >> gpurun_out/r2q_ab.txt; env $1 timeout 200 python tools/perf_nets.py $2 --reps 30 >> gpurun_out/r2q_ab.txt 2>&1; }
for b in 256 128 64 32; do run "FR_X=0" "--stages embed --emb-batch $b"; done
for b in 256 32; do run "FR_PDL=0" "--stages embed --emb-batch $b"; done
cat gpurun_out/r2q_ab.txt
timeout 600 python tools/run_bench_pipeline.py 20 > gpurun_out/r2q_pipeline.json 2> gpurun_out/r2q_pipeline.err
echo "pipeline rc=$?"; tail -c 300 gpurun_out/r2q_pipeline.err; cut -c1-500 gpurun_out/r2q_pipeline.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_detect_launches_b64.csv python tools/perf_nets.py --stages detect --det-batch 64 --reps 1 > /dev/null 2>&1
