#!/bin/bash
# round-2 GPU call J: symmetric-triangle Jacobi A/B (eig_diag + C2 bench), eigensolver parity tests, data path alone,
# ncu --set full of the tensor-core GEMM (tensor pipe %)
mkdir -p gpurun_out
timeout 300 python profiles/eig_diag.py > gpurun_out/j_eig_new.log 2>&1
GCCB200_LIB=build_variants/libgccb200_presym.so timeout 300 python profiles/eig_diag.py > gpurun_out/j_eig_presym.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity2.py tests/test_gpu_parity.py -q -m gpu -s -k "posenc or eig or engine_step or batch" > gpurun_out/j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/j_tests.log
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err
GCCB200_LIB=build_variants/libgccb200_presym.so timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/j_bench_presym.json 2> gpurun_out/j_bench_presym.err
timeout 300 python profiles/data_alone.py 4 > gpurun_out/j_data_alone.log 2>&1
timeout 300 python profiles/data_alone.py 1 >> gpurun_out/j_data_alone.log 2>&1
timeout 300 python profiles/train_alone.py >> gpurun_out/j_data_alone.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_kernel" -s 40 -c 6 \
    -o gpurun_out/j_prof_tc python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/j_ncu_tc.log 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/j_tests.log | cut -c1-250; tail -4 gpurun_out/j_eig_new.log | cut -c1-300; tail -4 gpurun_out/j_eig_presym.log | cut -c1-300
head -c 250 gpurun_out/j_bench.json; echo; head -c 250 gpurun_out/j_bench_presym.json; echo; cat gpurun_out/j_data_alone.log | grep -v Warn | tail -5
