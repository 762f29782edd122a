#!/bin/bash
# default bench on the committed tree (1 GPU) + the reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r2f4_bench.json 2> gpurun_out/r2f4_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2f4_bench.err
python - <<'P'
import json
d=json.loads([l for l in open("gpurun_out/r2f4_bench.json") if l.startswith("{")][-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["bound"], d["roofline"]["frac"], "traffic", d["roofline"]["traffic"])
for k,v in d["scans"].items():
    print(k, "q/s", round(v["value"]), "unknown", round(v["unknown_queries"]["value"]), "e2e", round(v["e2e"]["value"]), "kernel_ms", round(v["roofline"]["kernel_ms"],4), "frac", round(v["roofline"]["frac"],3))
p=d["pipeline"]
print("pipeline e2e", p["e2e"]["value"], "sync", p["e2e"]["synchronous_call"]["value"], p["e2e"]["parity_gate"]["identities_exact"], p["e2e"]["ms_per_batch_windows"])
for k in p:
    if k!="e2e": print(k, p[k].get("batch"), round(p[k]["ms"],3), "frac", round(p[k]["roofline"]["frac"],3))
P
