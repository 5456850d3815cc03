#!/bin/bash
# round-2 re-entry verification call (HEAD = gather / pooling / heads rewrites): GPU suite, smoke(), C2 bench with the CPU
# baseline, C4 bench, C4 launch list.  Short on purpose: 24 GPU-minutes were left in the round.
mkdir -p gpurun_out
P=gpurun_out/u
timeout 240 python -m pytest tests -q -m gpu > ${P}_tests.log 2>&1; echo "tests rc=$?" >> ${P}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${P}_smoke.log 2>&1
timeout 240 python bench.py --steps 200 --warmup 20 > ${P}_bench_c2.json 2> ${P}_bench_c2.err
timeout 150 python bench.py --config c4 --steps 40 --warmup 5 --no-cpu-baseline > ${P}_bench_c4_tc.json 2> ${P}_bench_c4_tc.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 2500 --csv --log-file ${P}_launches_c4.csv \
    python bench.py --config c4 --steps 6 --warmup 5 --no-cpu-baseline > ${P}_ncu_c4.log 2>&1
tail -3 ${P}_tests.log | cut -c1-300; tail -1 ${P}_smoke.log
for v in bench_c2 bench_c4_tc; do python - <<PY
import json
d=json.load(open("${P}_$v.json")); print("$v", round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), (d.get("cpu_baseline") or {}).get("value"))
PY
done
