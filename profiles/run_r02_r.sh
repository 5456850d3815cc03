#!/bin/bash
# round-2 GPU call R: config-5 A/B of the membership filter and the on-the-spot pool claim (rw_hops 256)
mkdir -p gpurun_out
for v in nobloom noclaim neither; do
GCCB200_LIB=build_variants/libgccb200_$v.so timeout 600 python profiles/sampler_sweep.py 24 200000000 32768 0 256 3 > gpurun_out/r_sweep_$v.json 2> gpurun_out/r_sweep_$v.err
done
for v in nobloom noclaim neither; do echo $v; grep -o "'rw_hops': [0-9]*, 'ms_per_launch_group': [0-9.]*, 'egonets_per_sec': [0-9.]*" gpurun_out/r_sweep_$v.err; done
