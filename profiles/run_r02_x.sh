#!/bin/bash
# dense eigensolver, third GPU call: idle-warp skipping, factor rows in shared memory (class A), 5 multisection rounds.
# Parity suite, phase split, pipeline bench over the remaining knobs.
mkdir -p gpurun_out
P=gpurun_out/x
timeout 300 python -m pytest tests -q -m gpu -x > ${P}_tests.log 2>&1; echo "tests rc=$?" >> ${P}_tests.log
timeout 100 compute-sanitizer --tool racecheck --racecheck-report all python profiles/dense_sanitize.py > ${P}_racecheck.log 2>&1; echo "rc=$?" >> ${P}_racecheck.log
timeout 100 python profiles/eig_dense_diag.py c2 3 > ${P}_dense_diag.log 2>&1
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 120 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > ${P}_bench_$name.json 2> ${P}_bench_$name.err
}
run d96 GCCB200_DENSE_MAX=96
run d96_nousm GCCB200_DENSE_MAX=96 GCCB200_DN_USM=0
run d96_a296 GCCB200_DENSE_MAX=96 GCCB200_DN_CAP_A=296
run d96_a592 GCCB200_DENSE_MAX=96 GCCB200_DN_CAP_A=592
run d144 GCCB200_DENSE_MAX=144
run d144_b148 GCCB200_DENSE_MAX=144 GCCB200_DN_CAP_B=148
run iter GCCB200_DENSE_MAX=0
tail -4 ${P}_tests.log | cut -c1-400; tail -3 ${P}_racecheck.log | cut -c1-300; tail -6 ${P}_dense_diag.log
for v in d96 d96_nousm d96_a296 d96_a592 d144 d144_b148 iter; do python - <<PY
import json
try:
    d=json.load(open("${P}_bench_$v.json")); print("%-16s" % "$v", round(d["value"]), round(d["ms_per_step"], 4), round(d["e2e"]["value"]), d["step_time"]["p50_ms"], d["step_time"]["p95_ms"], d["eigensolver"]["max_residual"], d["phases_ms"]["eigensolver_ms"])
except Exception as ex: print("$v failed", ex)
PY
done
