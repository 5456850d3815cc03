#!/bin/bash
# round-2 GPU call G: full GPU test suite, C2 + C4 bench lines, ncu --set full (with source) of the gather kernel and the n<=96 eigensolver kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -s > gpurun_out/g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/g_tests.log
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
timeout 600 python bench.py --config c4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/g_bench_c4.json 2> gpurun_out/g_bench_c4.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 2500 --csv --log-file gpurun_out/g_launches_c4.csv \
    python bench.py --config c4 --steps 6 --warmup 5 --no-cpu-baseline > gpurun_out/g_ncu_c4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gin_agg_cast_kernel|gin_bwd_dh_kernel" -s 10 -c 4 \
    -o gpurun_out/g_prof_gather python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/g_ncu_gather.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"posenc_chfsi_kernel" -s 4 -c 2 \
    -o gpurun_out/g_prof_eig python profiles/eig_diag.py > gpurun_out/g_ncu_eig.log 2>&1
grep -E "passed|failed|^FAILED|tensor-core" gpurun_out/g_tests.log | cut -c1-250; head -c 300 gpurun_out/g_bench.json; echo; head -c 300 gpurun_out/g_bench_c4.json
