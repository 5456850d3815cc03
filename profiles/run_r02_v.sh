#!/bin/bash
# dense tridiagonal eigensolver, first GPU call: parity suite, solver A/B on the same batches with phase counters,
# C2 bench with the dense solver on (default) and off
mkdir -p gpurun_out
P=gpurun_out/v
timeout 300 python -m pytest tests -q -m gpu -x > ${P}_tests.log 2>&1; echo "tests rc=$?" >> ${P}_tests.log
timeout 150 compute-sanitizer --tool racecheck --racecheck-report all python profiles/dense_sanitize.py > ${P}_racecheck.log 2>&1; echo "rc=$?" >> ${P}_racecheck.log
timeout 120 compute-sanitizer --tool memcheck python profiles/dense_sanitize.py > ${P}_memcheck.log 2>&1; echo "rc=$?" >> ${P}_memcheck.log
timeout 200 python profiles/eig_dense_diag.py c2 > ${P}_dense_diag.log 2>&1
timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > ${P}_bench_c2_dense.json 2> ${P}_bench_c2_dense.err
GCCB200_DENSE_MAX=0 timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > ${P}_bench_c2_iter.json 2> ${P}_bench_c2_iter.err
tail -4 ${P}_tests.log | cut -c1-400; tail -4 ${P}_racecheck.log | cut -c1-300; tail -3 ${P}_memcheck.log | cut -c1-300; tail -8 ${P}_dense_diag.log
for v in bench_c2_dense bench_c2_iter; do python - <<PY
import json
try:
    d=json.load(open("${P}_$v.json")); print("$v", round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["eigensolver"]["max_residual"], d["phases_ms"])
except Exception as ex: print("$v failed", ex)
PY
done
