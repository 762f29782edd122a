#!/bin/bash
# A/B of library builds (face-recognition-cpp-tensorrt_b200/lib_ab/<tag>): sustained search time per variant, interleaved, one box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
: > gpurun_out/r2_ab.log
for round in 1 2; do
for tag in default $AB_TAGS; do
  if [ "$tag" = default ]; then unset FR_B200_LIB; else export FR_B200_LIB=$PWD/face-recognition-cpp-tensorrt_b200/lib_ab/$tag/libfr_b200.so; fi
  timeout 300 python tools/ab_search.py ${AB_ROWS:-1250000} ${AB_CFG:-f8:unknown,f8:planted,f16:unknown} 2>&1 | tail -1 | tee -a gpurun_out/r2_ab.log
done
done
