#!/usr/bin/env python
"""GPU diagnostic: host enqueue time per engine step (is the pipeline host-bound?)."""
import contextlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gcc_b200  # noqa: F401  (sets CUDA_DEVICE_MAX_CONNECTIONS before CUDA starts)
import torch

import bench  # noqa: E402
from gcc_b200.contrastive.memory_moco import MemoryMoCo  # noqa: E402
from gcc_b200.datasets.graph_dataset import LoadBalanceGraphDataset  # noqa: E402
from gcc_b200.engine import PretrainEngine  # noqa: E402
from gcc_b200.models import GraphEncoder  # noqa: E402

cfg = bench.CONFIGS["c2"]
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
eng = PretrainEngine(ds, model, ema, contrast, moco=True, prefetch=int(sys.argv[1]) if len(sys.argv) > 1 else 2)
for _ in range(5):
    eng.step(lr=0.005)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(6):
    eng.step(lr=0.005)
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host enqueue, empty launch queue: %.3f ms/step" % (1e3 * (t1 - t0) / 6))
N = 40
t0 = time.perf_counter()
for _ in range(N):
    eng.step(lr=0.005)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue %.3f ms/step; with final sync %.3f ms/step" % (1e3 * (t1 - t0) / N, 1e3 * (t2 - t0) / N))
eng.timing, eng.timing_main = [], []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(N):
    eng.step(lr=0.005)
e1.record()
torch.cuda.synchronize()
tm, td = eng.timing_main, eng.timing
eng.timing = eng.timing_main = None
print("period %.3f ms; train span (after batch ready -> end of step) %.3f ms; gap between steps %.3f ms" % (
    e0.elapsed_time(e1) / N, sum(a.elapsed_time(b) for a, b in tm) / N,
    sum(tm[i][1].elapsed_time(tm[i + 1][0]) for i in range(N - 1)) / (N - 1)))
print("data: sampler %.3f eig %.3f ms per batch" % (sum(a.elapsed_time(b) for a, b, _ in td) / len(td),
                                                    sum(b.elapsed_time(c) for _, b, c in td) / len(td)))
# split: data preparation vs the training part (host side only)
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    eng.step(lr=0.005)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
