#!/bin/bash
# round-2 multi-GPU call: usage  bash profiles/run_r02_mg.sh N   (under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/mg${N}_smi.txt 2>&1
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_parity2.py -q -m gpu -s -k two_gpu > gpurun_out/mg2_replica_test.log 2>&1; echo "rc=$?" >> gpurun_out/mg2_replica_test.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/mg${N}_bench.json 2> gpurun_out/mg${N}_bench.err
tail -3 gpurun_out/mg2_replica_test.log 2>/dev/null; head -c 400 gpurun_out/mg${N}_bench.json; tail -3 gpurun_out/mg${N}_bench.err
