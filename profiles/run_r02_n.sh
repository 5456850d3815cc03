#!/bin/bash
# round-2 GPU call N: e2e stall diagnosis, scan-unroll sampler + 256-bit GEMM stores: tests, C2 / C4 bench, C5 point
mkdir -p gpurun_out
timeout 400 python profiles/e2e_diag.py 5 > gpurun_out/n_e2e_diag.log 2>&1
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_parity2.py -q -m gpu -k "tc or sampler or batch or wide or infonce or engine_step" > gpurun_out/n_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/n_tests.log
timeout 600 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err
timeout 600 python bench.py --config c4 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/n_bench_c4.json 2> gpurun_out/n_bench_c4.err
timeout 900 python profiles/sampler_sweep.py 24 200000000 32768 0 256 3 > gpurun_out/n_sweep_256.json 2> gpurun_out/n_sweep_256.err
grep -v Warn gpurun_out/n_e2e_diag.log | tail -24 | cut -c1-220
tail -3 gpurun_out/n_tests.log | cut -c1-300
for v in n_bench n_bench_c4; do python - <<PY
import json
d=json.load(open("gpurun_out/$v.json")); print("$v", round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["phases_ms"] if "phases_ms" in d else "")
PY
done
tail -c 500 gpurun_out/n_sweep_256.err
