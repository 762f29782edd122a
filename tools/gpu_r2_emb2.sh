#!/bin/bash
# embedder parity tests + forward times after the SE-gate weight prefetch
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_embedder_gpu.py -m gpu -q -x --timeout=600 2>&1 | tail -2
for b in 32 256 32 256; do
  timeout 300 python tools/perf_nets.py --stages embed --emb-batch $b --reps 200 2>&1 | tail -1 | cut -c1-140
done
