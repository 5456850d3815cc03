#!/usr/bin/env python
"""GPU diagnostic: steady-state throughput of the DATA path alone (sampler + eigensolver, S batches in
flight on the engine's data streams, no training part), to compare with the training part alone
(profiles/train_alone.py) and the full step: is the step bound by the sum of the two?

usage: data_alone.py [prefetch=4] [config=c2]"""
import contextlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gcc_b200  # noqa: F401  (sets CUDA_DEVICE_MAX_CONNECTIONS before CUDA starts)
import torch

import bench  # noqa: E402
from gcc_b200.contrastive.memory_moco import MemoryMoCo  # noqa: E402
from gcc_b200.datasets.graph_dataset import LoadBalanceGraphDataset  # noqa: E402
from gcc_b200.engine import PretrainEngine  # noqa: E402
from gcc_b200.models import GraphEncoder  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = bench.CONFIGS[sys.argv[2] if len(sys.argv) > 2 else "c2"]
dev = torch.device("cuda")
g = bench.make_graph_device(cfg, dev)
B, L, H, K = cfg["batch"], cfg["layers"], cfg["hidden"], cfg["K"]
ds = LoadBalanceGraphDataset(rw_hops=cfg["rw_hops"], restart_prob=0.8, dgl_graphs_file=g, batch_size=B, seed=0)


def mk():
    return GraphEncoder(positional_embedding_size=32, max_degree=512, degree_embedding_size=16, output_dim=H,
                        node_hidden_dim=H, num_layers=L, norm=True, gnn_model="gin", degree_input=True)


model, ema = mk(), mk()
ema.load_state_dict(model.state_dict())
model, ema = model.to(dev), ema.to(dev)
with contextlib.redirect_stdout(sys.stderr):
    contrast = MemoryMoCo(H, None, K, 0.07, use_softmax=True).to(dev)
eng = PretrainEngine(ds, model, ema, contrast, moco=True, prefetch=S)
for _ in range(6):
    eng.step(lr=0.005)
eng.wait_data_streams()
torch.cuda.synchronize()
main = torch.cuda.current_stream(dev)


def run(n):
    for _ in range(n):
        slot = eng.prepared % eng.depth
        eng._prepare(None)
        main.wait_event(eng.ready[slot])
        eng.consumed[slot].record(main)


run(8)
torch.cuda.synchronize()
N = 60
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
run(N)
e1.record()
torch.cuda.synchronize()
print("data path alone, %d batches in flight: %.3f ms/batch (%d ego-nets per batch)" % (S, e0.elapsed_time(e1) / N, 2 * B))
