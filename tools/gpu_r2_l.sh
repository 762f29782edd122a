#!/bin/bash
# session L: lean MMA issue loop in conv3x3_mt_kernel: parity, perf, ncu (tensor-pipe %) of stage-1 and stage-3 instances
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_embedder_gpu.py tests/test_pipeline_gpu.py -m gpu -q --timeout=300 > gpurun_out/r2l_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log; tail -15 gpurun_out/r2l_pytest.log | cut -c1-300
run() { echo "== $1 $2" >> gpurun_out/r2l_ab.txt; env $1 timeout 300 python tools/perf_nets.py $2 --reps 30 >> gpurun_out/r2l_ab.txt 2>&1; }
run "FR_X=0" "--stages embed --emb-batch 256"
run "FR_X=0" "--stages embed --emb-batch 32"
run "FR_X=0" "--stages embed --emb-batch 256 --arc-mode ir"
run "FR_X=0" "--stages e2e"
cat gpurun_out/r2l_ab.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_embed_launches_b256.csv python tools/perf_nets.py --stages embed --emb-batch 256 --reps 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_mt_kernel -s 144 -c 4 -f -o gpurun_out/r02_conv_mt_stage1_b256 python tools/perf_nets.py --stages embed --emb-batch 256 --reps 1 > gpurun_out/r2l_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_mt_kernel -s 165 -c 2 -f -o gpurun_out/r02_conv_mt_stage3_b256 python tools/perf_nets.py --stages embed --emb-batch 256 --reps 1 > gpurun_out/r2l_ncu2.log 2>&1
ls -la gpurun_out | tail -6
