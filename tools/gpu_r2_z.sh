#!/bin/bash
# session Z: source-level profile of the new e4m3 scan + re-rank (unmatched queries, 1.25 M rows); sustained 1-GPU bench at the 8-GPU shard size
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"cosine_topk_coarse|append_rerank" --launch-skip 6 -c 2 -f -o gpurun_out/r02_f8_unknown_1250k_after python tools/prof_search_unknown.py 1250000 f8 > gpurun_out/r2z_ncu_f8.log 2>&1
echo "ncu f8 rc=$?"
timeout 600 python bench.py --rows 1250000 --no-pipeline --no-cpu-baseline --no-ref-gpu --no-traffic > gpurun_out/r2z_bench_1250k.json 2> gpurun_out/r2z_bench_1250k.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2z_bench_1250k.err
python - <<'P'
import json
d=json.load(open("gpurun_out/r2z_bench_1250k.json"))
for k,v in d["scans"].items():
    print(k, "ms/step", round(v["ms_per_step"],4), "q/s", round(v["value"]), "unknown ms", round(v["unknown_queries"]["ms_per_step"],4), "e2e ms", round(v["e2e"]["ms_per_step"],4), "kernel_ms", round(v["roofline"]["kernel_ms"],4))
P
