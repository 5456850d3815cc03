#!/bin/bash
# round-2 GPU call B: all GPU tests, eigensolver schedule A/B, bench line, ncu launch list, ncu --set full of the sampler
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -s > gpurun_out/b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/b_tests.log
timeout 300 python profiles/eig_diag.py > gpurun_out/b_eig_new.log 2>&1
GCCB200_LIB=build_variants/libgccb200_old_sched.so timeout 300 python profiles/eig_diag.py > gpurun_out/b_eig_old.log 2>&1
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 2200 --csv --log-file gpurun_out/b_launches.csv \
    python bench.py --steps 10 --warmup 8 --no-cpu-baseline > gpurun_out/b_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rwr_walk|induce_fill|batch_offsets" -s 24 -c 6 \
    -o gpurun_out/b_prof_sampler python bench.py --steps 4 --warmup 4 --no-cpu-baseline > gpurun_out/b_ncu_sampler.log 2>&1
tail -4 gpurun_out/b_tests.log; grep "posenc ms\|cycles" gpurun_out/b_eig_new.log | head -8; grep "posenc ms" gpurun_out/b_eig_old.log | head -3; head -c 300 gpurun_out/b_bench.json
