#!/bin/bash
# fp8 append-epilogue experiments. Outputs: gpurun_out/f8_*.json
mkdir -p gpurun_out
T0=$(date +%s)
echo "== search tests"; timeout 600 python -m pytest tests/test_search_gpu.py -m gpu -q --timeout 300 -x 2>&1 | tail -8 | tee gpurun_out/f8_pytest.log
echo "t=$(( $(date +%s) - T0 ))s"
B="python bench.py --steps 20 --no-cpu-baseline --no-pipeline --no-fp8"
run() { name=$1; shift; echo "== $name"; timeout 300 "$@" 2>&1 | tail -1 > gpurun_out/f8_$name.json; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/f8_$name.json"))
    print("$name", "ms/step", round(d["ms_per_step"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "hbm_frac", round(d["roofline"]["frac"],3), "clk", d["clocks"]["sm_mhz"], d["parity"])
except Exception as e:
    print("$name FAILED", e, open("gpurun_out/f8_$name.json").read()[-400:])
PY
}
for R in 1250000 10000000; do
  run f8_planted_$R $B --rows $R --scan f8
  run f8_unknown_$R $B --rows $R --scan f8 --query-kind unknown
  FR_SEARCH_APPEND=0 run f8list_unknown_$R $B --rows $R --scan f8 --query-kind unknown
  FR_F8_EPS=1e-4 run f8_eps0_$R $B --rows $R --scan f8
  FR_SEARCH_APPEND=2 run f16app_planted_$R $B --rows $R
  FR_SEARCH_APPEND=2 run f16app_unknown_$R $B --rows $R --query-kind unknown
  run f16_unknown_$R $B --rows $R --query-kind unknown
done
echo "t=$(( $(date +%s) - T0 ))s"
