#!/bin/bash
# session X: per-kernel ncu metrics of every non-scan kernel (detector b16, embedder b32/b256, pipeline glue, exchange) on the final
# kernels (explicit metric list = few replay passes), and a full capture with source of the e4m3 scan + re-rank for unmatched queries
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python face-recognition-cpp-tensorrt_b200/build.py > /dev/null
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,lts__t_bytes.sum,launch__grid_size,launch__shared_mem_per_block_dynamic
for what in detect embed32 embed256 pipeline exchange; do
  timeout 400 ncu --metrics $M --clock-control none --nvtx --nvtx-include "prof/" --csv --page raw --log-file gpurun_out/r02_ncu_raw_$what.csv python tools/prof_target.py $what > gpurun_out/r2x_ncu_$what.log 2>&1
  echo "ncu $what rc=$? lines=$(wc -l < gpurun_out/r02_ncu_raw_$what.csv)"
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"cosine_topk_coarse|append_rerank" --launch-skip 6 -c 2 -f -o gpurun_out/r02_f8_unknown_1250k python tools/prof_search_unknown.py 1250000 f8 > gpurun_out/r2x_ncu_f8.log 2>&1
echo "ncu f8 rc=$?"; ls -la gpurun_out/*.ncu-rep | tail -3
