#!/bin/bash
# search stream with a copy stream and three batches in flight: tests + e2e vs device-resident at the 8-GPU shard size
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_search_gpu.py -m gpu -q -x --timeout=600 -k "stream or push or exchange" > gpurun_out/r2s2_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2s2_pytest.log | cut -c1-200
timeout 600 python bench.py --rows 1250000 --no-pipeline --no-cpu-baseline --no-ref-gpu --no-traffic > gpurun_out/r2s2_bench_1250k.json 2> gpurun_out/r2s2_bench_1250k.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2s2_bench_1250k.err
python - <<'P'
import json
d=json.loads([l for l in open("gpurun_out/r2s2_bench_1250k.json") if l.startswith("{")][-1])
for k,v in d["scans"].items():
    print(k, "ms/step", round(v["ms_per_step"],4), "unknown ms", round(v["unknown_queries"]["ms_per_step"],4), "e2e ms", round(v["e2e"]["ms_per_step"],4), "kernel_ms", round(v["roofline"]["kernel_ms"],4))
P
