#!/bin/bash
# session K: 256-bit epilogue stores, coalesced stem stores, faster SE gate, detector side chains: parity, perf, launch lists, pipeline bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_detector_gpu.py tests/test_embedder_gpu.py tests/test_pipeline_gpu.py tests/test_dropin_cpp.py -m gpu -q --timeout=300 > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log; tail -25 gpurun_out/r2k_pytest.log | cut -c1-300
run() { echo "== $1 $2" >> gpurun_out/r2k_ab.txt; env $1 timeout 300 python tools/perf_nets.py $2 --reps 30 >> gpurun_out/r2k_ab.txt 2>&1; }
run "FR_X=0" "--stages embed --emb-batch 256"
run "FR_STEM_TC=0" "--stages embed --emb-batch 256"
run "FR_X=0" "--stages embed --emb-batch 32"
run "FR_X=0" "--stages detect,e2e"
run "FR_DET_LANES=0" "--stages detect"
run "FR_PIPE_SUB=32" "--stages e2e"
run "FR_PIPE_SUB=64" "--stages e2e"
cat gpurun_out/r2k_ab.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_detect_launches_b16.csv python tools/perf_nets.py --stages detect --reps 1 > /dev/null 2>&1
for b in 32 256; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_embed_launches_b$b.csv python tools/perf_nets.py --stages embed --emb-batch $b --reps 1 > /dev/null 2>&1
done
ls -la gpurun_out | tail -6
