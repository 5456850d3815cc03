#!/usr/bin/env python
"""GPU diagnostic: the training part of the step alone (no sampler / eigensolver in flight), on the
whole GPU.  usage: train_alone.py [config=c2]"""
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

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
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
eng = PretrainEngine(ds, model, ema, contrast, moco=True, prefetch=2)
for _ in range(4):
    eng.step(lr=0.005)
torch.cuda.synchronize()
N = 30
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(eng.train_stream):
    for _ in range(3):
        eng._step(0.005, None, True)
    from gcc_b200 import _lib
    l0 = _lib.get().gccb_launch_count()
    e0.record()
    for _ in range(N):
        eng._step(0.005, None, True)
    e1.record()
    launches = _lib.get().gccb_launch_count() - l0
torch.cuda.synchronize()
print("train part alone: %.3f ms/step, %d library launches/step" % (e0.elapsed_time(e1) / N, launches / N))
