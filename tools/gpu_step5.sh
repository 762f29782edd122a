#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8 | tee gpurun_out/s5_pytest.log
echo "t=$(( $(date +%s) - T0 ))s"
for B in 32 256; do timeout 200 python tools/perf_nets.py --stages embed --emb-batch $B 2>&1 | tail -1 | tee -a gpurun_out/s5_perf_embed.json; done
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== default bench"; timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/s5_bench_default.json; python - <<'PY'
import json
d=json.load(open("gpurun_out/s5_bench_default.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","clocks")}, d["e2e"], d["roofline"], d["roofline_tensor"]["frac"], d["other_scan"], d["cpu_baseline"]["value"], d.get("ref_gpu"))
print(d["pipeline"]["e2e"]["value"], d["pipeline"]["detect"]["ms"], d["pipeline"]["embed"]["ms"])
PY
echo "t=$(( $(date +%s) - T0 ))s"
