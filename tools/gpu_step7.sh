#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
echo "== search + pipeline tests"; timeout 600 python -m pytest tests/test_search_gpu.py tests/test_pipeline_gpu.py -m gpu -q --timeout 300 2>&1 | tail -6 | tee gpurun_out/s7_pytest.log
echo "t=$(( $(date +%s) - T0 ))s"
B="python bench.py --steps 30 --no-cpu-baseline --no-pipeline --no-alt-scan"
for R in 1250000 10000000; do timeout 300 $B --rows $R 2>&1 | tail -1 > gpurun_out/s7_f8_$R.json; python - <<PY
import json
d=json.load(open("gpurun_out/s7_f8_$R.json"))
print($R, "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "hbm_frac", round(d["roofline"]["frac"],3), "clk", d["clocks"]["sm_mhz"], "launches", d["gpu_launches"], d["parity"])
PY
done
echo "t=$(( $(date +%s) - T0 ))s"
