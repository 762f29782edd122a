#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
nvidia-smi -L | wc -l
run() { N=$1; EX=$2; shift 2; echo "== bench $N GPUs exchange=$EX $*"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) bench.py --gpus $N --steps 30 --warmup 5 --exchange $EX "$@" 2>&1 | grep '^{' | tail -1 > gpurun_out/g${N}_bench_${EX}${TAG}.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/g${N}_bench_${EX}${TAG}.json"))
    o=d.get("other_scan") or {}
    print("N=$N $EX", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "hbm_frac", round(d["roofline"]["frac"],3), "clk", d["clocks"], "| other:", o.get("scan"), o.get("value"))
except Exception as e:
    print("FAILED", e)
PY
}
TAG="" run 8 p2p
TAG="" run 4 p2p
TAG="" run 8 nccl --no-alt-scan
TAG="_unknown" run 8 p2p --query-kind unknown --no-alt-scan
echo "t=$(( $(date +%s) - T0 ))s"
