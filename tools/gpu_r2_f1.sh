#!/bin/bash
# final validation 1: the whole GPU suite, then compute-sanitizer memcheck over the e4m3 search tests (new epilogue / re-rank)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r2f1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f1_pytest.log; tail -5 gpurun_out/r2f1_pytest.log | cut -c1-300
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 \
  python -m pytest tests/test_search_gpu.py -m gpu -q -x -k "fp8_scan_copy_topk or fp8_scan_unknown or search_stream" -p no:cacheprovider > gpurun_out/r02_sanitizer_memcheck_f8.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/r02_sanitizer_memcheck_f8.log
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02_sanitizer_memcheck_f8.log | tail -4
