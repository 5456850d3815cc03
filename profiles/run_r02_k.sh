#!/bin/bash
# round-2 GPU call K: first-Ritz-solve / deflated-start-block variants of the eigensolver: eig_diag + C2 bench each
mkdir -p gpurun_out
for v in s0def s0 s1def; do
  GCCB200_LIB=build_variants/libgccb200_$v.so timeout 300 python profiles/eig_diag.py > gpurun_out/k_eig_$v.log 2>&1
  GCCB200_LIB=build_variants/libgccb200_$v.so timeout 600 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/k_bench_$v.json 2> gpurun_out/k_bench_$v.err
done
timeout 600 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/k_bench_cur.json 2> gpurun_out/k_bench_cur.err
GCCB200_LIB=build_variants/libgccb200_s0def.so timeout 900 python -m pytest tests/test_gpu_parity2.py tests/test_gpu_parity.py -q -m gpu -s -k "posenc or eig or engine_step or batch" > gpurun_out/k_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/k_tests.log
GCCB200_LIB=build_variants/libgccb200_s0def.so timeout 300 python profiles/data_alone.py 4 > gpurun_out/k_data_alone.log 2>&1
for v in s0def s0 s1def cur; do echo $v; head -c 230 gpurun_out/k_bench_$v.json | tail -c 110; echo; grep -E "posenc ms|\(64,96\]: cycles" gpurun_out/k_eig_$v.log 2>/dev/null | tail -2 | cut -c1-200; done
grep -E "passed|failed|^FAILED" gpurun_out/k_tests.log | cut -c1-250; grep "data path" gpurun_out/k_data_alone.log
