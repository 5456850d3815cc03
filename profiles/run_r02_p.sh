#!/bin/bash
# round-2 GPU call P: config 5 after the pool-capacity fix: full sweep, A/B of the scan unroll and the CTA-wide hub rows, ncu split
mkdir -p gpurun_out
timeout 900 python profiles/sampler_sweep.py 24 200000000 32768 0 64,128,256,512 3 > gpurun_out/p_sweep.json 2> gpurun_out/p_sweep.err
GCCB200_LIB=build_variants/libgccb200_u1.so timeout 600 python profiles/sampler_sweep.py 24 200000000 32768 0 256 3 > gpurun_out/p_sweep_u1.json 2> gpurun_out/p_sweep_u1.err
GCCB200_LIB=build_variants/libgccb200_h1.so timeout 600 python profiles/sampler_sweep.py 24 200000000 32768 0 256 3 > gpurun_out/p_sweep_h1.json 2> gpurun_out/p_sweep_h1.err
timeout 900 ncu --set full --clock-control none -k regex:"rwr_walk_unique_kernel|induce_fill_kernel" -s 4 -c 2 -o gpurun_out/p_prof_c5 \
    python profiles/sampler_sweep.py 24 200000000 32768 0 256 1 > gpurun_out/p_ncu_c5.json 2> gpurun_out/p_ncu_c5.err
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py -q -m gpu -k "sampler or batch" > gpurun_out/p_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/p_tests.log
for v in p_sweep p_sweep_u1 p_sweep_h1; do grep -o "'rw_hops': [0-9]*, 'ms_per_launch_group': [0-9.]*, 'egonets_per_sec': [0-9.]*" gpurun_out/$v.err; done
tail -2 gpurun_out/p_tests.log
