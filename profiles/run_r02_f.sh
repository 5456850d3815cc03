#!/bin/bash
# round-2 GPU call F: wide-GIN gradient diagnostic, eigensolver A/B (packed Jacobi records), ncu --set full of the C4 hot kernels
mkdir -p gpurun_out
timeout 600 python profiles/wide_grad_check.py 256 > gpurun_out/f_wide256.log 2>&1
timeout 600 python profiles/wide_grad_check.py 128 > gpurun_out/f_wide128.log 2>&1
timeout 300 python profiles/eig_diag.py > gpurun_out/f_eig_new.log 2>&1
GCCB200_LIB=build_variants/libgccb200_pre_rec.so timeout 300 python profiles/eig_diag.py > gpurun_out/f_eig_prerec.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gin_agg_cast_kernel|tc_gemm_kernel|cast_bf16_kernel|gin_bn_tail_kernel|gin_bwd_dh_kernel" -s 120 -c 16 \
    -o gpurun_out/f_prof_c4 python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f_ncu_c4.log 2>&1
cat gpurun_out/f_wide256.log | tail -40; tail -12 gpurun_out/f_wide128.log; grep "posenc ms\|cycles" gpurun_out/f_eig_new.log | head -6; grep "posenc ms\|cycles" gpurun_out/f_eig_prerec.log | head -6
