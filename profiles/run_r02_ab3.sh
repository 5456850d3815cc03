#!/bin/bash
# last GPU call of round 2: the ILP / branch-free rewrite of the dense solver's loops against the previous build
# (build_variants/libgccb200_prev.so = the library of commit "tests: fuzz of the dense eigensolver ...")
mkdir -p gpurun_out
P=gpurun_out/ab3
timeout 120 python -m pytest tests -q -m gpu -x -k "eigensolver or posenc or dense" > ${P}_tests.log 2>&1; echo "tests rc=$?" >> ${P}_tests.log
timeout 60 python bench.py --steps 150 --warmup 20 --no-cpu-baseline > ${P}_bench_new.json 2> ${P}_bench_new.err
GCCB200_LIB=build_variants/libgccb200_prev.so timeout 60 python bench.py --steps 150 --warmup 20 --no-cpu-baseline > ${P}_bench_prev.json 2> ${P}_bench_prev.err
timeout 40 python profiles/eig_dense_diag.py c2 2 > ${P}_dense_diag.log 2>&1
tail -3 ${P}_tests.log | cut -c1-300; tail -3 ${P}_dense_diag.log | cut -c1-400
for v in new prev; do python - <<PY
import json
try:
    d=json.load(open("${P}_bench_$v.json")); print("$v", round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["step_time"]["p50_ms"])
except Exception as ex: print("$v failed", ex)
PY
done
