#!/bin/bash
# A/B of library builds on the end-to-end pipeline bench (in-flight and synchronous faces/s)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
: > gpurun_out/r2_abp.log
for round in 1 2; do
for tag in default $AB_TAGS; do
  if [ "$tag" = default ]; then unset FR_B200_LIB; else export FR_B200_LIB=$PWD/face-recognition-cpp-tensorrt_b200/lib_ab/$tag/libfr_b200.so; fi
  timeout 400 python tools/run_bench_pipeline.py 20 2>/dev/null | tail -1 | python -c "
import json,sys
p=json.loads(sys.stdin.read()); e=p['e2e']
print('$tag', 'in-flight', round(e['value']), 'ms', round(e['ms_per_batch'],3), 'sync', round(e['synchronous_call']['value']), 'detect64', round(p.get('detect',{}).get('ms',0),3), 'embed256', round(p['embed_b256']['ms'],3), e['parity_gate']['identities_exact'])
" | tee -a gpurun_out/r2_abp.log
done
done
