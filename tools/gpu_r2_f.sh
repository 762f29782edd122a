#!/bin/bash
# re-entry session F: full GPU parity suite, default bench (both arms), launch lists of the nets
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log; tail -4 gpurun_out/r2f_pytest.log
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2f_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/r2f_ref.json 2> gpurun_out/r2f_ref.err
echo "ref rc=$?"; cut -c1-300 gpurun_out/r2f_ref.json
timeout 600 python tools/perf_nets.py > gpurun_out/r2f_perf_nets.txt 2>&1
tail -30 gpurun_out/r2f_perf_nets.txt
