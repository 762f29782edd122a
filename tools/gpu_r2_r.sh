#!/bin/bash
# session R: the whole GPU suite + the full bench (both arms) on the current tree
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r2r_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log; tail -6 gpurun_out/r2r_pytest.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2r_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/r2r_ref.json 2> gpurun_out/r2r_ref.err
echo "ref rc=$?"; cut -c1-200 gpurun_out/r2r_ref.json
