#!/bin/bash
# round-2 final evidence call: full GPU suite, smoke(), bench lines (C2 with CPU baseline, reference arm, C4 tc / simt,
# E2E mode), launch lists, ncu --set full of the GEMM / eigensolver / sampler, CUPTI timelines, alone-runs
mkdir -p gpurun_out
P=gpurun_out/f
timeout 1500 python -m pytest tests -q -m gpu > ${P}_tests.log 2>&1; echo "tests rc=$?" >> ${P}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${P}_smoke.log 2>&1
timeout 900 python bench.py --steps 200 --warmup 20 > ${P}_bench_c2.json 2> ${P}_bench_c2.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${P}_bench_ref.json 2> ${P}_bench_ref.err
timeout 600 python bench.py --config c4 --steps 40 --warmup 5 --no-cpu-baseline > ${P}_bench_c4_tc.json 2> ${P}_bench_c4_tc.err
GCCB200_TC=0 timeout 600 python bench.py --config c4 --steps 20 --warmup 5 --no-cpu-baseline > ${P}_bench_c4_simt.json 2> ${P}_bench_c4_simt.err
timeout 600 python bench.py --mode e2e --steps 200 --warmup 20 --no-cpu-baseline > ${P}_bench_e2emode.json 2> ${P}_bench_e2emode.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file ${P}_launches_c2.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > ${P}_ncu_c2.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 2500 --csv --log-file ${P}_launches_c4.csv \
    python bench.py --config c4 --steps 6 --warmup 5 --no-cpu-baseline > ${P}_ncu_c4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_kernel" -s 40 -c 4 -o ${P}_prof_tc \
    python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline > ${P}_ncu_tc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"posenc_chfsi_kernel|posenc_chfsi_cluster_kernel|posenc_jacobi_kernel" -s 5 -c 5 -o ${P}_prof_eig \
    python profiles/eig_diag.py > ${P}_ncu_eig.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"rwr_walk_unique_kernel|induce_fill_kernel|batch_offsets_kernel" -s 12 -c 6 -o ${P}_prof_sampler \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${P}_ncu_sampler.log 2>&1
timeout 300 python profiles/eig_diag.py > ${P}_eig_diag.log 2>&1
timeout 300 python profiles/timeline.py 4 ${P}_timeline_c2.json.gz c2 > ${P}_tl_c2.log 2>&1 && python profiles/timeline_read.py ${P}_timeline_c2.json.gz x > ${P}_tl_c2_summary.txt 2>&1
timeout 300 python profiles/timeline.py 4 ${P}_timeline_c4.json.gz c4 > ${P}_tl_c4.log 2>&1 && python profiles/timeline_read.py ${P}_timeline_c4.json.gz x > ${P}_tl_c4_summary.txt 2>&1
{ timeout 300 python profiles/data_alone.py 4; timeout 300 python profiles/data_alone.py 1; timeout 300 python profiles/train_alone.py; } 2>&1 | grep -E "data path|train part" > ${P}_alone.log
timeout 900 python profiles/sampler_sweep.py 24 200000000 32768 0 64,128,256,512 3 > ${P}_sweep_c5.json 2> ${P}_sweep_c5.err
timeout 900 ncu --set full --clock-control none -k regex:"rwr_walk_unique_kernel|induce_fill_kernel" -s 4 -c 2 -o ${P}_prof_c5 \
    python profiles/sampler_sweep.py 24 200000000 32768 0 256 1 > ${P}_ncu_c5.json 2> ${P}_ncu_c5.err
tail -3 ${P}_tests.log | cut -c1-300; cat ${P}_smoke.log | tail -2
grep -o "'rw_hops': [0-9]*, 'ms_per_launch_group': [0-9.]*, 'egonets_per_sec': [0-9.]*" ${P}_sweep_c5.err
for v in bench_c2 bench_ref bench_c4_tc bench_c4_simt bench_e2emode; do head -c 300 ${P}_$v.json | cut -c1-300; echo; done
cat ${P}_alone.log
