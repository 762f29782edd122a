#!/bin/bash
# session H: fused detector kernels (stem block, dw+pw small, split decode/NMS, merged SSH GEMM, dual c16, heads epilogue),
# exact-pooling SE: parity, then perf + launch lists
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_detector_gpu.py tests/test_embedder_gpu.py tests/test_pipeline_gpu.py tests/test_dropin_cpp.py -m gpu -q --timeout=300 > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log; tail -25 gpurun_out/r2h_pytest.log | cut -c1-300
timeout 300 python tools/perf_nets.py --reps 30 > gpurun_out/r2h_perf.txt 2>&1
timeout 300 python tools/perf_nets.py --stages embed --emb-batch 256 --reps 30 >> gpurun_out/r2h_perf.txt 2>&1
timeout 300 python tools/perf_nets.py --stages detect --det-batch 64 --reps 30 >> gpurun_out/r2h_perf.txt 2>&1
cat gpurun_out/r2h_perf.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_detect_launches_b16.csv python tools/perf_nets.py --stages detect --reps 1 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_embed_launches_b256.csv python tools/perf_nets.py --stages embed --emb-batch 256 --reps 1 > /dev/null 2>&1
ls -la gpurun_out | tail -8
