#!/usr/bin/env python
"""GPU diagnostic: where does an end-to-end step (host seeds -> H2D -> PretrainEngine.step -> stats D2H) stall?
Replays bench.py's e2e loop several times with host timestamps around every part of a step and prints the
steps that took more than 3x the median, with the part that was slow."""
import contextlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gcc_b200  # noqa: F401
import numpy as np
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
eng = PretrainEngine(ds, model, ema, contrast, moco=True, prefetch=4)
WARM = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for _ in range(WARM):
    eng.step(lr=0.005)
torch.cuda.synchronize()
cdf_host = ds.graph.cdf.cpu().numpy()
rs = np.random.RandomState(1)
N, REP, LAG = 20, 6, 2
for rep in range(REP):
    host_seeds = torch.from_numpy(np.searchsorted(cdf_host, rs.random_sample((N + 2, B)), side="right")
                                  .clip(max=len(cdf_host) - 1).astype(np.int64)).pin_memory()
    loss_ring = torch.zeros(4, 4, dtype=torch.float32).pin_memory()
    done = [torch.cuda.Event() for _ in range(4)]
    NSEED = 12
    seeds_ring = [torch.zeros(B, dtype=torch.int64, device=dev) for _ in range(NSEED)]
    if rep == 0 or len(sys.argv) > 2:
        for i in range(2):
            seeds_ring[i].copy_(host_seeds[i], non_blocking=True)
            eng.step(lr=0.005, seeds=seeds_ring[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    T = np.zeros((N, 5))
    e0.record()
    for i in range(N):
        T[i, 0] = time.perf_counter()
        seeds_ring[(i + 2) % NSEED].copy_(host_seeds[2 + i], non_blocking=True)
        T[i, 1] = time.perf_counter()
        eng.step(lr=0.005, seeds=seeds_ring[(i + 2) % NSEED])
        T[i, 2] = time.perf_counter()
        loss_ring[i & 3].copy_(eng.stats, non_blocking=True)
        done[i & 3].record()
        T[i, 3] = time.perf_counter()
        if i >= LAG:
            done[(i - LAG) & 3].synchronize()
            float(loss_ring[(i - LAG) & 3][0])
        T[i, 4] = time.perf_counter()
    torch.cuda.synchronize()
    eng.wait_data_streams()
    e1.record()
    torch.cuda.synchronize()
    tot = e0.elapsed_time(e1)
    d = np.diff(T, axis=1) * 1e3
    step_ms = d.sum(axis=1)
    med = np.median(step_ms)
    print("rep %d: %d steps in %.1f ms (%.0f subgraphs/s); host per step median %.2f ms; parts median [copy %.3f step %.3f d2h %.3f sync %.3f]"
          % (rep, N, tot, 2 * B * N / tot * 1e3, med, *np.median(d, axis=0)))
    for i in np.flatnonzero(step_ms > 3 * med):
        print("   slow step %d: %.2f ms = copy %.2f | eng.step %.2f | d2h+record %.2f | sync %.2f" % (i, step_ms[i], *d[i]))
