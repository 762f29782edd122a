#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
nvidia-smi -L
echo "== search tests"; timeout 600 python -m pytest tests/test_search_gpu.py -m gpu -q --timeout 300 -x 2>&1 | tail -5 | tee gpurun_out/g2_pytest.log
echo "t=$(( $(date +%s) - T0 ))s"
N=${NGPU:-2}
for EX in p2p nccl; do
echo "== bench $N GPUs exchange=$EX"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --exchange $EX 2>&1 | grep '^{' | tail -1 | tee gpurun_out/g${N}_bench_$EX.json | cut -c1-900
done
echo "== bench $N GPUs unknown queries"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --query-kind unknown --no-alt-scan 2>&1 | grep '^{' | tail -1 | tee gpurun_out/g${N}_bench_unknown.json | cut -c1-400
echo "== reference arm under torchrun"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 3 --warmup 1 --no-pipeline 2>&1 | grep '^{' | tail -1 | cut -c1-300
echo "t=$(( $(date +%s) - T0 ))s"
