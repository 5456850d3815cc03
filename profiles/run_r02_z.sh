#!/bin/bash
# round-2 final numbers (dense eigensolver for n <= 96, empty dense classes not launched): GPU suite, smoke(), C2 bench with
# the CPU baseline, C4 bench, E2E-mode bench, C2 launch list
mkdir -p gpurun_out
P=gpurun_out/z
timeout 300 python -m pytest tests -q -m gpu > ${P}_tests.log 2>&1; echo "tests rc=$?" >> ${P}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${P}_smoke.log 2>&1
timeout 240 python bench.py --steps 200 --warmup 20 > ${P}_bench_c2.json 2> ${P}_bench_c2.err
timeout 150 python bench.py --config c4 --steps 40 --warmup 5 --no-cpu-baseline > ${P}_bench_c4_tc.json 2> ${P}_bench_c4_tc.err
timeout 100 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --mode e2e > ${P}_bench_c2_e2emode.json 2> ${P}_bench_c2_e2emode.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file ${P}_launches_c2.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > ${P}_ncu_c2.log 2>&1
tail -3 ${P}_tests.log | cut -c1-300; tail -1 ${P}_smoke.log
for v in bench_c2 bench_c4_tc bench_c2_e2emode; do python - <<PY
import json
try:
    d=json.load(open("${P}_$v.json")); print("$v", round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), (d.get("cpu_baseline") or {}).get("value"), d.get("gpu_launches_per_step"))
except Exception as ex: print("$v failed", ex)
PY
done
