#!/bin/bash
# session J: split SE (gate + streaming apply), tensor-core stem, stride-2 convs on the MT kernel: parity, A/B, launch lists
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_detector_gpu.py tests/test_embedder_gpu.py tests/test_pipeline_gpu.py tests/test_dropin_cpp.py -m gpu -q --timeout=300 > gpurun_out/r2j_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log; tail -25 gpurun_out/r2j_pytest.log | cut -c1-300
run() { echo "== $1 batch=$2" >> gpurun_out/r2j_ab.txt; env $1 timeout 300 python tools/perf_nets.py --stages embed --emb-batch $2 --reps 30 >> gpurun_out/r2j_ab.txt 2>&1; }
for b in 256 32; do
  run "FR_X=0" $b
  run "FR_STEM_TC=0" $b
  run "FR_MT_S2=0" $b
done
run "FR_X=0" 128
run "FR_X=0" 64
timeout 300 python tools/perf_nets.py --stages detect,e2e --reps 30 >> gpurun_out/r2j_ab.txt 2>&1
cat gpurun_out/r2j_ab.txt
for b in 32 256; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_embed_launches_b$b.csv python tools/perf_nets.py --stages embed --emb-batch $b --reps 1 > /dev/null 2>&1
done
ls -la gpurun_out | tail -6
