#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_search_gpu.py -m gpu -q --timeout=600 -k "structured or outside_fp16 or roster or search_stream or fused_push" > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log; tail -4 gpurun_out/r2e_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2e_bench.err
