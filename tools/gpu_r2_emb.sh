#!/bin/bash
# A/B: pair kernel admitted on partially filled grids (FR_PAIR_MINFILL) at small embedder batches
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
: > gpurun_out/r2_emb_ab.log
for round in 1 2; do
for mf in 1.0 0.75 0.5; do
for b in 32 64; do
  echo "== FR_PAIR_MINFILL=$mf batch $b" | tee -a gpurun_out/r2_emb_ab.log
  FR_PAIR_MINFILL=$mf timeout 300 python tools/perf_nets.py --stages embed --emb-batch $b --reps 200 2>&1 | tail -1 | tee -a gpurun_out/r2_emb_ab.log
done
done
done
