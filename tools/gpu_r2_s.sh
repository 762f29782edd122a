#!/bin/bash
# session S (N GPUs): bench.py under torchrun, both arms
N=${1:-2}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 1200 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2s_bench_${N}gpu.json 2> gpurun_out/r2s_bench_${N}gpu.err
echo "bench$N rc=$?"; tail -c 600 gpurun_out/r2s_bench_${N}gpu.err | tail -5; cut -c1-300 gpurun_out/r2s_bench_${N}gpu.json
