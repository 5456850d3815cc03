#!/bin/bash
# round-2 GPU call L: first-sweep split + 16-warp GEMM epilogue: tests, C2 / C4 bench, occupancy-pad variant, GEMM ncu
mkdir -p gpurun_out
timeout 300 python profiles/eig_diag.py > gpurun_out/l_eig.log 2>&1
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity2.py tests/test_gpu_parity.py -q -m gpu -s > gpurun_out/l_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/l_tests.log
timeout 600 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err
GCCB200_LIB=build_variants/libgccb200_pad2.so timeout 600 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/l_bench_pad2.json 2> gpurun_out/l_bench_pad2.err
timeout 600 python bench.py --config c4 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/l_bench_c4.json 2> gpurun_out/l_bench_c4.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_kernel" -s 40 -c 4 \
    -o gpurun_out/l_prof_tc python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/l_ncu_tc.log 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/l_tests.log | cut -c1-250; grep -E "posenc ms|cycles" gpurun_out/l_eig.log | cut -c1-200 | tail -6
for v in l_bench l_bench_pad2 l_bench_c4; do head -c 230 gpurun_out/$v.json | tail -c 110; echo; done
