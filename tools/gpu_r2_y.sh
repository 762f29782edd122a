#!/bin/bash
# session Y: vote-descent append epilogue + per-list / exact-leader re-rank: search parity suite, launch lists (unknown queries, 1.25 M rows)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_search_gpu.py tests/test_search_reference_lib.py -m gpu -q -x --timeout=600 > gpurun_out/r2y_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2y_pytest.log; tail -6 gpurun_out/r2y_pytest.log | cut -c1-300
for sc in f8 f16; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02y_search_unknown_launches_${sc}_1250k.csv python tools/prof_search_unknown.py 1250000 $sc > gpurun_out/r2y_$sc.log 2>&1
tail -2 gpurun_out/r2y_$sc.log
done
python - <<'P'
import csv,collections
for sc in ("f8","f16"):
    rows=[r for r in csv.reader(open(f"gpurun_out/r02y_search_unknown_launches_{sc}_1250k.csv")) if len(r)>5]
    h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
    d=collections.defaultdict(list)
    for r in rows[1:]:
        try: d[r[ki][:36]].append(float(r[vi].replace(",","")))
        except ValueError: pass
    for k,v in d.items(): print(sc,k,len(v),round(sum(v)/len(v)/1000,1),"us")
P
