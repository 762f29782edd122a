#!/bin/bash
# 2-GPU session: sharded search (fused push + lagged merge), NCCL baseline, pipeline replicas
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c_bench2.json 2> gpurun_out/r2c_bench2.err
echo "bench2 rc=$?"; tail -c 800 gpurun_out/r2c_bench2.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-lag --no-pipeline --no-alt-scan > gpurun_out/r2c_bench2_nolag.json 2> gpurun_out/r2c_bench2_nolag.err
echo "nolag rc=$?"; tail -c 400 gpurun_out/r2c_bench2_nolag.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --exchange nccl --no-pipeline --no-alt-scan > gpurun_out/r2c_bench2_nccl.json 2> gpurun_out/r2c_bench2_nccl.err
echo "nccl rc=$?"; tail -c 400 gpurun_out/r2c_bench2_nccl.err
