#!/bin/bash
# round-2 closing evidence call (final build: dense eigensolver for n <= 96 by default): GPU suite, smoke(), C2 bench with the
# CPU baseline, C4 bench, two run-ahead depths, C2 launch list, ncu --set full of the dense kernel, CUPTI timeline.
mkdir -p gpurun_out
P=gpurun_out/y
timeout 300 python -m pytest tests -q -m gpu > ${P}_tests.log 2>&1; echo "tests rc=$?" >> ${P}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${P}_smoke.log 2>&1
timeout 100 compute-sanitizer --tool racecheck --racecheck-report all python profiles/dense_sanitize.py > ${P}_racecheck.log 2>&1; echo "rc=$?" >> ${P}_racecheck.log
timeout 240 python bench.py --steps 200 --warmup 20 > ${P}_bench_c2.json 2> ${P}_bench_c2.err
timeout 150 python bench.py --config c4 --steps 40 --warmup 5 --no-cpu-baseline > ${P}_bench_c4_tc.json 2> ${P}_bench_c4_tc.err
timeout 100 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --prefetch 3 > ${P}_bench_c2_s3.json 2> ${P}_bench_c2_s3.err
timeout 100 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --prefetch 6 > ${P}_bench_c2_s6.json 2> ${P}_bench_c2_s6.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file ${P}_launches_c2.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > ${P}_ncu_c2.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"posenc_dense_kernel" -s 2 -c 2 -o ${P}_prof_dense \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${P}_ncu_dense.log 2>&1
timeout 120 python profiles/timeline.py 4 ${P}_timeline_c2.json.gz c2 > ${P}_tl_c2.log 2>&1 && python profiles/timeline_read.py ${P}_timeline_c2.json.gz x > ${P}_tl_c2_summary.txt 2>&1
tail -3 ${P}_tests.log | cut -c1-300; tail -1 ${P}_smoke.log; tail -2 ${P}_racecheck.log | cut -c1-200
for v in bench_c2 bench_c4_tc bench_c2_s3 bench_c2_s6; do python - <<PY
import json
try:
    d=json.load(open("${P}_$v.json")); print("$v", round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), (d.get("cpu_baseline") or {}).get("value"))
except Exception as ex: print("$v failed", ex)
PY
done
