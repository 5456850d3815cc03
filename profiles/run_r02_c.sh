#!/bin/bash
# round-2 GPU call C: tensor-core GIN path parity + config-4 lines (bf16 tcgen05 vs fp32 SIMT)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity2.py tests/test_gpu_tc.py -q -m gpu -s -k "wide or tc_gemm or c2_batch_posenc or every_size" > gpurun_out/c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c_tests.log
timeout 600 python bench.py --config c4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c_bench_c4_tc.json 2> gpurun_out/c_bench_c4_tc.err
GCCB200_TC=0 timeout 600 python bench.py --config c4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c_bench_c4_simt.json 2> gpurun_out/c_bench_c4_simt.err
tail -15 gpurun_out/c_tests.log | cut -c1-300; head -c 400 gpurun_out/c_bench_c4_tc.json; echo; head -c 400 gpurun_out/c_bench_c4_simt.json; tail -3 gpurun_out/c_bench_c4_tc.err
