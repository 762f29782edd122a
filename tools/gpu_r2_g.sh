#!/bin/bash
# session G: persistent multi-tile conv (conv3x3_mt_kernel) + fused SE: parity, A/B against conv_gemm_kernel, launch lists
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_embedder_gpu.py tests/test_pipeline_gpu.py -m gpu -q --timeout=300 > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log; tail -15 gpurun_out/r2g_pytest.log
for mode in ir_se ir; do
for b in 32 256; do
  for mt in 1 0; do
    echo "== arc=$mode batch=$b FR_CONV_MT=$mt" >> gpurun_out/r2g_ab.txt
    FR_CONV_MT=$mt timeout 300 python tools/perf_nets.py --stages embed --emb-batch $b --reps 30 --arc-mode $mode >> gpurun_out/r2g_ab.txt 2>&1
  done
done
done
for f in 128,2 128,1 64,2 64,1; do
  echo "== force $f batch=256" >> gpurun_out/r2g_ab.txt
  FR_MT_FORCE=$f timeout 300 python tools/perf_nets.py --stages embed --emb-batch 256 --reps 30 >> gpurun_out/r2g_ab.txt 2>&1
  echo "== force $f batch=32" >> gpurun_out/r2g_ab.txt
  FR_MT_FORCE=$f timeout 300 python tools/perf_nets.py --stages embed --emb-batch 32 --reps 30 >> gpurun_out/r2g_ab.txt 2>&1
done
cat gpurun_out/r2g_ab.txt
for b in 32 256; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_embed_launches_b$b.csv python tools/perf_nets.py --stages embed --emb-batch $b --reps 1 > /dev/null 2>&1
done
ls -la gpurun_out
