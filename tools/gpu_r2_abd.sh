#!/bin/bash
# A/B of library builds on the detector (device-resident forward), interleaved on one box; then the detector parity tests on the default build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
: > gpurun_out/r2_abd.log
for round in 1 2; do
for tag in default $AB_TAGS; do
  if [ "$tag" = default ]; then unset FR_B200_LIB; else export FR_B200_LIB=$PWD/face-recognition-cpp-tensorrt_b200/lib_ab/$tag/libfr_b200.so; fi
  timeout 300 python tools/ab_detect.py 2>&1 | tail -1 | tee -a gpurun_out/r2_abd.log
done
done
unset FR_B200_LIB
timeout 600 python -m pytest tests/test_detector_gpu.py -m gpu -q -x --timeout=600 2>&1 | tail -2
