#!/bin/bash
# round-end style check: build is prebuilt in-tree; GPU tests, smoke, default bench, reference arm
mkdir -p gpurun_out
T0=$(date +%s)
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -4 | tee gpurun_out/final_pytest.log
echo "t=$(( $(date +%s) - T0 ))s"
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/final_bench.json; python - <<'PY'
import json
d=json.load(open("gpurun_out/final_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","steps","gpu_launches","clocks")}, d["e2e"], {k:d["roofline"][k] for k in ("achieved","frac","kernel_ms","traffic")}, d["roofline_tensor"]["frac"])
print("other", d["other_scan"]["value"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "ref_gpu", d.get("ref_gpu",{}).get("queries_per_s_at_sample"))
print("pipeline", d["pipeline"]["e2e"]["value"], d["pipeline"].get("e2e_two_in_flight"), d["pipeline"]["detect"]["ms"], d["pipeline"]["embed"]["ms"])
PY
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-300
echo "t=$(( $(date +%s) - T0 ))s"
