#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
echo "== lifecycle + search tests"; timeout 600 python -m pytest tests/test_search_gpu.py -m gpu -q --timeout 300 2>&1 | tail -6 | tee gpurun_out/s6_pytest.log
echo "t=$(( $(date +%s) - T0 ))s"
echo "== pipeline"; timeout 600 python - <<'PY' 2>&1 | tail -3 | tee gpurun_out/s6_pipeline.json
import json, sys
sys.path.insert(0, "."); sys.path.insert(0, "face-recognition-cpp-tensorrt_b200")
from tools import bench_pipeline as bp
print(json.dumps(bp.run_gpu(0)))
PY
echo "t=$(( $(date +%s) - T0 ))s"
echo "== ncu full (final scan kernel, e4m3, 10M)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cosine_topk_coarse -s 4 -c 1 -f -o gpurun_out/coarse_f8_final_10M \
    python bench.py --steps 1 --warmup 3 --ramp-s 0 --no-cpu-baseline --no-pipeline --no-alt-scan --no-graph > gpurun_out/ncu_full_f8_final.log 2>&1
tail -1 gpurun_out/ncu_full_f8_final.log | cut -c1-150
echo "t=$(( $(date +%s) - T0 ))s"
