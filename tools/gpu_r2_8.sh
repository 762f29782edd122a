#!/bin/bash
# 8-GPU search bench (no pipeline section): the north-star configuration
N=${1:-8}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-pipeline --no-cpu-baseline --no-ref-gpu --no-traffic > gpurun_out/r2_8_bench_${N}gpu.json 2> gpurun_out/r2_8_bench_${N}gpu.err
echo "bench$N rc=$?"; tail -c 400 gpurun_out/r2_8_bench_${N}gpu.err | tail -4
python - <<P
import json
d=json.loads([l for l in open("gpurun_out/r2_8_bench_${N}gpu.json") if l.startswith("{")][-1])
print("headline", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
for k,v in d["scans"].items():
    print(k, "ms/step", round(v["ms_per_step"],4), "q/s", round(v["value"]), "unknown q/s", round(v["unknown_queries"]["value"]), "e2e q/s", round(v["e2e"]["value"]), "kernel_ms", round(v["roofline"]["kernel_ms"],4), "frac", round(v["roofline"]["frac"],3), v["parity"])
P
