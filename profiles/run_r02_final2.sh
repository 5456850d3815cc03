#!/bin/bash
# round-2 closing evidence call (after the gather / pooling kernel rewrites): GPU suite, smoke(), C2 bench with the CPU
# baseline, C4 bench, launch lists and CUPTI timelines of both, ncu of the gather kernels, alone-runs at C4
mkdir -p gpurun_out
P=gpurun_out/g2
timeout 1500 python -m pytest tests -q -m gpu > ${P}_tests.log 2>&1; echo "tests rc=$?" >> ${P}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${P}_smoke.log 2>&1
timeout 900 python bench.py --steps 200 --warmup 20 > ${P}_bench_c2.json 2> ${P}_bench_c2.err
timeout 600 python bench.py --config c4 --steps 40 --warmup 5 --no-cpu-baseline > ${P}_bench_c4_tc.json 2> ${P}_bench_c4_tc.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file ${P}_launches_c2.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > ${P}_ncu_c2.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 2500 --csv --log-file ${P}_launches_c4.csv \
    python bench.py --config c4 --steps 6 --warmup 5 --no-cpu-baseline > ${P}_ncu_c4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gin_agg_cast_kernel|gin_bwd_dh_kernel|gin_pool_kernel|gin_pool_predict_kernel" -s 10 -c 6 -o ${P}_prof_gather \
    python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline > ${P}_ncu_gather.log 2>&1
timeout 300 python profiles/timeline.py 4 ${P}_timeline_c2.json.gz c2 > ${P}_tl_c2.log 2>&1 && python profiles/timeline_read.py ${P}_timeline_c2.json.gz x > ${P}_tl_c2_summary.txt 2>&1
timeout 300 python profiles/timeline.py 4 ${P}_timeline_c4.json.gz c4 > ${P}_tl_c4.log 2>&1 && python profiles/timeline_read.py ${P}_timeline_c4.json.gz x > ${P}_tl_c4_summary.txt 2>&1
{ timeout 300 python profiles/data_alone.py 4 c4; timeout 300 python profiles/train_alone.py c4; } 2>&1 | grep -E "data path|train part" > ${P}_alone_c4.log
tail -3 ${P}_tests.log | cut -c1-300; tail -1 ${P}_smoke.log
for v in bench_c2 bench_c4_tc; do python - <<PY
import json
d=json.load(open("${P}_$v.json")); print("$v", round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), (d.get("cpu_baseline") or {}).get("value"))
PY
done
cat ${P}_alone_c4.log
