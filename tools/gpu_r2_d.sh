#!/bin/bash
# hardware session D: new GPU tests, bench (whole-block graphs), ncu raw metrics of all non-scan kernels, sanitizer
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log; tail -4 gpurun_out/r2d_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2d_bench.err
timeout 1200 ncu --set full --clock-control none --nvtx --nvtx-include "prof/" --csv --page raw --log-file gpurun_out/r02_nets_ncu_raw.csv python tools/prof_target.py > gpurun_out/r2d_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2d_ncu.log; ls -la gpurun_out/r02_nets_ncu_raw.csv
bash tools/sanitize.sh
