#!/bin/bash
# round-2 GPU call T: pooling kernels (128-bit segmented sum, 8 graphs per CTA in the prediction heads): GPU suite, C4 / C2 bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_tests.log
timeout 600 python bench.py --config c4 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/t_bench_c4.json 2> gpurun_out/t_bench_c4.err
timeout 600 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err
timeout 300 python profiles/train_alone.py c4 2>&1 | grep -E "train part" > gpurun_out/t_alone_c4.log
timeout 300 python profiles/timeline.py 4 gpurun_out/t_timeline_c4.json.gz c4 > gpurun_out/t_tl_c4.log 2>&1 && python profiles/timeline_read.py gpurun_out/t_timeline_c4.json.gz x > gpurun_out/t_tl_c4_summary.txt 2>&1
tail -3 gpurun_out/t_tests.log | cut -c1-300
for v in t_bench_c4 t_bench; do python - <<PY
import json
d=json.load(open("gpurun_out/$v.json")); print("$v", round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]))
PY
done
cat gpurun_out/t_alone_c4.log; sed -n '/^kernel/,/^step window/p' gpurun_out/t_tl_c4_summary.txt | head -30
