#!/bin/bash
# round-2 GPU call Q: sampler (single-look rows, membership filter): tests, config-5 sweep, C2 / C4 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py -q -m gpu -k "sampler or batch or engine_step or prefetch" > gpurun_out/q_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/q_tests.log
timeout 900 python profiles/sampler_sweep.py 24 200000000 32768 0 64,128,256,512 3 > gpurun_out/q_sweep.json 2> gpurun_out/q_sweep.err
timeout 600 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
timeout 600 python bench.py --config c4 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/q_bench_c4.json 2> gpurun_out/q_bench_c4.err
tail -2 gpurun_out/q_tests.log
grep -o "'rw_hops': [0-9]*, 'ms_per_launch_group': [0-9.]*, 'egonets_per_sec': [0-9.]*" gpurun_out/q_sweep.err
for v in q_bench q_bench_c4; do python - <<PY
import json
d=json.load(open("gpurun_out/$v.json")); print("$v", round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["phases_ms"]["sampler_ms"])
PY
done
