#!/bin/bash
# round-2 GPU call A: tcgen05 GEMM test in its own process (a trap must not poison the rest), all GPU tests,
# eigensolver schedule A/B, one short bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_tc.py -q -m gpu > gpurun_out/a_tc.log 2>&1; echo "tc rc=$?" >> gpurun_out/a_tc.log
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_tc.py -s > gpurun_out/a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/a_tests.log
timeout 300 python profiles/eig_diag.py > gpurun_out/a_eig_new.log 2>&1
GCCB200_LIB=build_variants/libgccb200_old_sched.so timeout 300 python profiles/eig_diag.py > gpurun_out/a_eig_old.log 2>&1
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
tail -3 gpurun_out/a_tc.log; tail -5 gpurun_out/a_tests.log; grep "posenc ms\|cycles" gpurun_out/a_eig_new.log | head -8; grep "posenc ms\|cycles" gpurun_out/a_eig_old.log | head -8; head -c 600 gpurun_out/a_bench.json
