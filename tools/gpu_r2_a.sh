#!/bin/bash
# round 2, first hardware session: GPU parity suite, the new bench (both arms), nothing else
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc > gpurun_out/r2a_host.txt; free -g >> gpurun_out/r2a_host.txt
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2a_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 --no-pipeline > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err
echo "ref rc=$?"; cut -c1-400 gpurun_out/r2a_ref.json
