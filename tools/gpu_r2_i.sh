#!/bin/bash
# session I: int32 exact SE pooling, L1-direct fused dw+pw, fused dw+pw GEMM (dwpw_gemm_kernel): parity, perf, A/B, ncu of the SE kernel + MT conv
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_detector_gpu.py tests/test_embedder_gpu.py tests/test_pipeline_gpu.py tests/test_dropin_cpp.py -m gpu -q --timeout=300 > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log; tail -25 gpurun_out/r2i_pytest.log | cut -c1-300
timeout 300 python tools/perf_nets.py --reps 30 > gpurun_out/r2i_perf.txt 2>&1
timeout 300 python tools/perf_nets.py --stages embed --emb-batch 256 --reps 30 >> gpurun_out/r2i_perf.txt 2>&1
echo "== FR_DET_FUSED_GEMM=0" >> gpurun_out/r2i_perf.txt
FR_DET_FUSED_GEMM=0 timeout 300 python tools/perf_nets.py --stages detect --reps 30 >> gpurun_out/r2i_perf.txt 2>&1
cat gpurun_out/r2i_perf.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_detect_launches_b16.csv python tools/perf_nets.py --stages detect --reps 1 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_embed_launches_b256.csv python tools/perf_nets.py --stages embed --emb-batch 256 --reps 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:se_gate_apply -s 34 -c 1 -f -o gpurun_out/r02_se_gate_apply_b256 python tools/perf_nets.py --stages embed --emb-batch 256 --reps 1 > gpurun_out/r2i_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_mt_kernel -s 109 -c 2 -f -o gpurun_out/r02_conv_mt_b256 python tools/perf_nets.py --stages embed --emb-batch 256 --reps 1 > gpurun_out/r2i_ncu2.log 2>&1
tail -3 gpurun_out/r2i_ncu1.log gpurun_out/r2i_ncu2.log
ls -la gpurun_out | tail -12
