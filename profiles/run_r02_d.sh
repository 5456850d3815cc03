#!/bin/bash
# round-2 GPU call D: tensor-core parity (GIN + InfoNCE), panel Gram-Schmidt A/B, C2 and C4 bench lines, C4 launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity2.py tests/test_gpu_tc.py -q -m gpu -s -k "wide or tc_gemm or c2_batch_posenc or every_size or infonce" > gpurun_out/d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/d_tests.log
timeout 300 python profiles/eig_diag.py > gpurun_out/d_eig_new.log 2>&1
GCCB200_LIB=build_variants/libgccb200_gs_scalar.so timeout 300 python profiles/eig_diag.py > gpurun_out/d_eig_scalar.log 2>&1
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
timeout 600 python bench.py --config c4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/d_bench_c4_tc.json 2> gpurun_out/d_bench_c4_tc.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 2500 --csv --log-file gpurun_out/d_launches_c4.csv \
    python bench.py --config c4 --steps 6 --warmup 5 --no-cpu-baseline > gpurun_out/d_ncu_c4.log 2>&1
grep -E "passed|failed|tensor-core|tc_gemm M" gpurun_out/d_tests.log | cut -c1-300; grep "posenc ms\|cycles" gpurun_out/d_eig_new.log | head -7; grep "posenc ms\|cycles" gpurun_out/d_eig_scalar.log | head -7; head -c 300 gpurun_out/d_bench.json; echo; head -c 300 gpurun_out/d_bench_c4_tc.json
