#!/bin/bash
# round-2 GPU call I: in-situ kernel timelines (CUPTI) of C2 and C4 steps; eigensolver A/B of the final build
mkdir -p gpurun_out
timeout 600 python profiles/timeline.py 4 gpurun_out/i_timeline_c4.json.gz c4 > gpurun_out/i_tl_c4.log 2>&1
timeout 300 python profiles/timeline_read.py gpurun_out/i_timeline_c4.json.gz > gpurun_out/i_tl_c4_summary.txt 2>&1
timeout 600 python profiles/timeline.py 4 gpurun_out/i_timeline_c2.json.gz c2 > gpurun_out/i_tl_c2.log 2>&1
timeout 300 python profiles/timeline_read.py gpurun_out/i_timeline_c2.json.gz x > gpurun_out/i_tl_c2_summary.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "engine_step" > gpurun_out/i_tests.log 2>&1
head -50 gpurun_out/i_tl_c4_summary.txt; tail -3 gpurun_out/i_tests.log
