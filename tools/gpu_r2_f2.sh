#!/bin/bash
# final validation 2: the accumulation test on the e4m3 copy, then the full default bench (1 GPU)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_search_gpu.py -m gpu -q -x --timeout=600 -k "accumulation or fp8_scan_copy_topk" > gpurun_out/r2f2_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2f2_pytest.log | cut -c1-200
timeout 1200 python bench.py > gpurun_out/r2f2_bench.json 2> gpurun_out/r2f2_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2f2_bench.err
python - <<'P'
import json
d=json.loads([l for l in open("gpurun_out/r2f2_bench.json") if l.startswith("{")][-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["bound"], d["roofline"]["frac"], "traffic", d["roofline"]["traffic"])
for k,v in d["scans"].items():
    print(k, "q/s", round(v["value"]), "unknown", round(v["unknown_queries"]["value"]), "e2e", round(v["e2e"]["value"]), "kernel_ms", round(v["roofline"]["kernel_ms"],4), "frac", round(v["roofline"]["frac"],3))
p=d["pipeline"]
print("pipeline e2e", p["e2e"]["value"], "sync", p["e2e"]["synchronous_call"]["value"], p["e2e"]["parity_gate"]["identities_exact"])
for k in p:
    if k!="e2e": print(k, p[k].get("batch"), round(p[k]["ms"],3), "frac", round(p[k]["roofline"]["frac"],3))
print("cpu", d["cpu_baseline"]["value"], "ref_gpu", d.get("ref_gpu",{}).get("value"))
P
