#!/bin/bash
# session N: pair kernel with split last round: parity, A/B; then the full bench (both arms)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_embedder_gpu.py tests/test_pipeline_gpu.py -m gpu -q --timeout=200 > gpurun_out/r2n_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log; tail -15 gpurun_out/r2n_pytest.log | cut -c1-300
run() { echo "== $1 $2" >> gpurun_out/r2n_ab.txt; env $1 timeout 200 python tools/perf_nets.py $2 --reps 30 >> gpurun_out/r2n_ab.txt 2>&1; }
run "FR_X=0" "--stages embed --emb-batch 256"
run "FR_PAIR_SPLIT=0" "--stages embed --emb-batch 256"
run "FR_X=0" "--stages embed --emb-batch 128"
run "FR_X=0" "--stages embed --emb-batch 64"
run "FR_X=0" "--stages embed --emb-batch 32"
run "FR_X=0" "--stages e2e"
cat gpurun_out/r2n_ab.txt
timeout 900 python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2n_bench.err
