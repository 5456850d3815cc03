#!/bin/bash
# round-2 GPU call H: full GPU test suite, C2 + C4 bench lines, C4 launch list (after the epilogue / gather / elementwise rework)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -s > gpurun_out/h_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/h_tests.log
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
timeout 600 python bench.py --config c4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/h_bench_c4.json 2> gpurun_out/h_bench_c4.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 2500 --csv --log-file gpurun_out/h_launches_c4.csv \
    python bench.py --config c4 --steps 6 --warmup 5 --no-cpu-baseline > gpurun_out/h_ncu_c4.log 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/h_tests.log | cut -c1-250; head -c 300 gpurun_out/h_bench.json; echo; head -c 300 gpurun_out/h_bench_c4.json
