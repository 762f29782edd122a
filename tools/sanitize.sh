#!/bin/bash
# compute-sanitizer target: memcheck + racecheck over a bounded subset of the GPU suite (the sanitizer slows kernels 10-50x).
# Logs land in gpurun_out/; copy the summaries to profiles/.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
SEL='config1 or adversarial or fused_push or missing_peer or fp8_saturating or starts_empty or exact_scan_of or test_crop_resize'
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 20 \
    python -m pytest tests/test_search_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x -k "$SEL" -p no:cacheprovider \
    > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/r02_sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r02_sanitizer_$tool.log | tail -4
done
# the e4m3 top-1 path (append scan, per-list re-rank with its filters and compactions) on a gallery small enough for racecheck
timeout 300 compute-sanitizer --tool racecheck --print-limit 10 python tools/racecheck_f8_small.py > gpurun_out/r02_racecheck_f8_small.log 2>&1
grep -E "RACECHECK SUMMARY|flagged|top1" gpurun_out/r02_racecheck_f8_small.log
