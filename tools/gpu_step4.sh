#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -6 | tee gpurun_out/s4_pytest.log
echo "t=$(( $(date +%s) - T0 ))s"
B="python bench.py --steps 20 --no-cpu-baseline --no-pipeline"
run() { name=$1; shift; echo "== $name"; timeout 300 "$@" 2>&1 | tail -1 > gpurun_out/s4_$name.json; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s4_$name.json"))
    o=d.get("other_scan") or {}
    print("$name", "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "hbm_frac", round(d["roofline"]["frac"],3), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "launches", d["gpu_launches"], "| other:", o.get("scan"), o.get("ms_per_step"), o.get("kernel_ms"))
except Exception as e:
    print("$name FAILED", e, open("gpurun_out/s4_$name.json").read()[-400:])
PY
}
run f8_1250000 $B --rows 1250000
run f8_unknown_1250000 $B --rows 1250000 --query-kind unknown --no-alt-scan
echo "t=$(( $(date +%s) - T0 ))s"
echo "== default bench"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/s4_bench_default.json | cut -c1-1200
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee gpurun_out/s4_bench_reference.json | cut -c1-600
echo "t=$(( $(date +%s) - T0 ))s"
