#!/bin/bash
# round-2 GPU call O: CTA-wide hub rows in the sampler: tests, C5 point, C2 bench; short driver-like runs with e2e host timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py -q -m gpu -k "sampler or batch or engine_step or prefetch" > gpurun_out/o_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/o_tests.log
timeout 900 python profiles/sampler_sweep.py 24 200000000 32768 0 256 3 > gpurun_out/o_sweep_256.json 2> gpurun_out/o_sweep_256.err
timeout 600 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err
for i in 1 2 3 4 5; do timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/o_bench_short$i.json 2> gpurun_out/o_bench_short$i.err; done
tail -3 gpurun_out/o_tests.log | cut -c1-300
tail -c 500 gpurun_out/o_sweep_256.err
for v in o_bench o_bench_short1 o_bench_short2 o_bench_short3 o_bench_short4 o_bench_short5; do python - <<PY
import json
d=json.load(open("gpurun_out/$v.json")); print("$v", round(d["value"]), round(d["e2e"]["value"]), d["phases_ms"]["sampler_ms"], {k:(round(v,2) if isinstance(v,float) else v) for k,v in d["e2e"]["host_step_ms"].items() if k!="note"})
PY
done
