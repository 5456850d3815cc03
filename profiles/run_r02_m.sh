#!/bin/bash
# round-2 GPU call M: finetune tests, C5 sampler sweep (unclipped budgets) + ncu DRAM bytes, host cost, short driver-like bench runs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_finetune.py -q -m gpu -x > gpurun_out/m_tests_ft.log 2>&1; echo "tests rc=$?" >> gpurun_out/m_tests_ft.log
timeout 300 python profiles/host_cost.py 4 > gpurun_out/m_host_cost.log 2>&1
for i in 1 2 3; do timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/m_bench_short$i.json 2> gpurun_out/m_bench_short$i.err; done
timeout 1200 python profiles/sampler_sweep.py 24 200000000 32768 0 64,128,256,512 3 > gpurun_out/m_sweep_unclipped.json 2> gpurun_out/m_sweep_unclipped.err
timeout 900 ncu --set full --clock-control none -k regex:"rwr_walk_unique_kernel|induce_fill_kernel" -s 4 -c 2 -o gpurun_out/m_prof_c5 \
    python profiles/sampler_sweep.py 24 200000000 32768 0 256 1 > gpurun_out/m_ncu_c5.json 2> gpurun_out/m_ncu_c5.err
tail -3 gpurun_out/m_tests_ft.log | cut -c1-300; grep -E "host enqueue|period" gpurun_out/m_host_cost.log
for i in 1 2 3; do python - <<PY
import json
d=json.load(open("gpurun_out/m_bench_short$i.json")); print("short", round(d["value"]), round(d["e2e"]["value"]), d["step_time"]["p50_ms"], d["step_time"]["max_ms"])
PY
done
tail -c 600 gpurun_out/m_sweep_unclipped.err; head -c 400 gpurun_out/m_sweep_unclipped.json
