#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 30 --warmup 5 2>&1 | grep '^{' | tail -1 > gpurun_out/q${N}_bench_p2p.json
python - <<PY
import json
d=json.load(open("gpurun_out/q${N}_bench_p2p.json"))
o=d.get("other_scan") or {}
print("N=$N", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "launches", d["gpu_launches"], d["parity"], "| f16:", o.get("value"))
PY
