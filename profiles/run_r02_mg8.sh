#!/bin/bash
# round-2 multi-GPU call on an 8-GPU box: N=8 and N=4 bench lines (weak scaling), C2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/mg8_smi.txt 2>&1
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N \
    bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/mg8_bench_n$N.json 2> gpurun_out/mg8_bench_n$N.err
done
timeout 300 python bench.py --gpus 1 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/mg8_bench_n1.json 2> gpurun_out/mg8_bench_n1.err
for N in 8 4 1; do head -c 260 gpurun_out/mg8_bench_n$N.json | tail -c 150; echo; done
