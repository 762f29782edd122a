#!/bin/bash
# session M: CTA-pair conv kernel (conv3x3_pair_kernel): parity, A/B, launch list, ncu of one stage-3 instance
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_embedder_gpu.py tests/test_pipeline_gpu.py -m gpu -q --timeout=200 > gpurun_out/r2m_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log; tail -15 gpurun_out/r2m_pytest.log | cut -c1-300
run() { echo "== $1 $2" >> gpurun_out/r2m_ab.txt; env $1 timeout 200 python tools/perf_nets.py $2 --reps 30 >> gpurun_out/r2m_ab.txt 2>&1; }
run "FR_X=0" "--stages embed --emb-batch 256"
run "FR_PAIR=0" "--stages embed --emb-batch 256"
run "FR_PAIR_BN=128" "--stages embed --emb-batch 256"
run "FR_PAIR_BN=256" "--stages embed --emb-batch 256"
run "FR_X=0" "--stages embed --emb-batch 128"
run "FR_X=0" "--stages embed --emb-batch 32"
run "FR_X=0" "--stages embed --emb-batch 256 --arc-mode ir"
run "FR_X=0" "--stages e2e"
cat gpurun_out/r2m_ab.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_embed_launches_b256.csv python tools/perf_nets.py --stages embed --emb-batch 256 --reps 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_pair_kernel -s 110 -c 2 -f -o gpurun_out/r02_conv_pair_stage3_b256 python tools/perf_nets.py --stages embed --emb-batch 256 --reps 1 > gpurun_out/r2m_ncu1.log 2>&1
ls -la gpurun_out | tail -6
