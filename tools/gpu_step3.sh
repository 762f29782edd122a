#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
echo "== hbm read probe"; timeout 120 tools/_bin/hbm_read 2>&1 | tee gpurun_out/hbm_read_probe.jsonl
echo "== tests"; timeout 600 python -m pytest tests/test_search_gpu.py tests/test_pipeline_gpu.py -m gpu -q --timeout 300 -x 2>&1 | tail -6 | tee gpurun_out/s3_pytest.log
echo "t=$(( $(date +%s) - T0 ))s"
B="python bench.py --steps 20 --no-cpu-baseline --no-pipeline"
run() { name=$1; shift; echo "== $name"; timeout 300 "$@" 2>&1 | tail -1 > gpurun_out/s3_$name.json; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s3_$name.json"))
    o=d.get("other_scan") or {}
    print("$name", "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "hbm_frac", round(d["roofline"]["frac"],3), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "launches", d["gpu_launches"], "| other:", o.get("scan"), o.get("ms_per_step"), o.get("kernel_ms"))
except Exception as e:
    print("$name FAILED", e, open("gpurun_out/s3_$name.json").read()[-400:])
PY
}
for R in 1250000 10000000; do
  run f8_$R $B --rows $R
  run f8_unknown_$R $B --rows $R --query-kind unknown --no-alt-scan
done
echo "t=$(( $(date +%s) - T0 ))s"
echo "== ncu launch lists"
for R in 1250000 10000000; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_f8_$R.csv \
    python bench.py --rows $R --steps 2 --warmup 3 --ramp-s 0 --no-cpu-baseline --no-pipeline --no-alt-scan --no-graph > gpurun_out/ncu_launches_f8_$R.log 2>&1
tail -1 gpurun_out/ncu_launches_f8_$R.log | cut -c1-200
done
echo "t=$(( $(date +%s) - T0 ))s"
