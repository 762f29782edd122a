#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
echo "== search tests (tiled e4m3 copy, append epilogue on both copies)"; timeout 600 python -m pytest tests/test_search_gpu.py tests/test_pipeline_gpu.py tests/test_search_reference_lib.py -m gpu -q --timeout 300 -x 2>&1 | tail -8 | tee gpurun_out/f8b_pytest.log
echo "t=$(( $(date +%s) - T0 ))s"
B="python bench.py --steps 20 --no-cpu-baseline --no-pipeline --no-alt-scan"
run() { name=$1; shift; echo "== $name"; timeout 300 "$@" 2>&1 | tail -1 > gpurun_out/f8b_$name.json; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/f8b_$name.json"))
    print("$name", "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "hbm_frac", round(d["roofline"]["frac"],3), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["parity"])
except Exception as e:
    print("$name FAILED", e, open("gpurun_out/f8b_$name.json").read()[-400:])
PY
}
for R in 1250000 10000000; do
  run tiled_$R $B --rows $R
  FR_F8_TILED=0 run rowmajor_$R $B --rows $R
  run tiled_unknown_$R $B --rows $R --query-kind unknown
  run tiled_again_$R $B --rows $R
done
echo "t=$(( $(date +%s) - T0 ))s"
