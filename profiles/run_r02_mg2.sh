#!/bin/bash
# 2-GPU check of the final build: replica identity test and the N = 2 bench line
mkdir -p gpurun_out
P=gpurun_out/mg2
timeout 200 python -m pytest tests -q -m gpu -k "two_gpu" > ${P}_tests.log 2>&1; echo "tests rc=$?" >> ${P}_tests.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 20 > ${P}_bench_c2_n2.json 2> ${P}_bench_c2_n2.err
tail -3 ${P}_tests.log | cut -c1-300
python - <<PY
import json
try:
    d=json.load(open("${P}_bench_c2_n2.json")); print("n2", round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d.get("rank_skew"))
except Exception as ex: print("n2 failed", ex)
PY
