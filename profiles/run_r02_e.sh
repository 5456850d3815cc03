#!/bin/bash
# round-2 GPU call E: scheduling A/B on C2 (run-ahead depth, eigensolver grid caps), training part alone, E2E-mode line
mkdir -p gpurun_out
for s in 2 3 5; do timeout 300 python bench.py --steps 150 --warmup 15 --no-cpu-baseline --prefetch $s > gpurun_out/e_bench_s$s.json 2> gpurun_out/e_bench_s$s.err; done
for v in mid1_296 mid1_222; do GCCB200_LIB=build_variants/libgccb200_$v.so timeout 300 python bench.py --steps 150 --warmup 15 --no-cpu-baseline > gpurun_out/e_bench_$v.json 2> gpurun_out/e_bench_$v.err; done
timeout 300 python profiles/train_alone.py > gpurun_out/e_train_alone.log 2>&1
timeout 300 python bench.py --steps 150 --warmup 15 --no-cpu-baseline --mode e2e > gpurun_out/e_bench_e2emode.json 2> gpurun_out/e_bench_e2emode.err
for f in gpurun_out/e_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['value']), round(d['ms_per_step'],3), 'p50', round(d['step_time']['p50_ms'],3), 'p95', round(d['step_time']['p95_ms'],3), 'e2e', round(d['e2e']['value']), 'samp/eig', round(d['phases_ms']['sampler_ms'],2), round(d['phases_ms']['eigensolver_ms'],2))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done; tail -2 gpurun_out/e_train_alone.log
