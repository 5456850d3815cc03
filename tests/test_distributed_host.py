"""CPU, world_size = 2, gloo: the host-side protocol of the one-collective data-parallel step
(gcc_b200/parallel.py): sample-id sharding, payload packing, all-gather, rank-ordered gradient sum
and queue order.  (The arithmetic on the gathered buffer runs in libgccb200 on the GPU; here it is
mirrored with torch to validate ordering and indexing.)"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gcc_b200.parallel import StepExchange, first_sample_id
    B, d, n_live, K = 4, 8, 37, 32
    torch.manual_seed(100 + rank)
    queue = torch.zeros(K, d)
    index = 0
    params = torch.zeros(n_live)
    ids = []
    for step in range(3):
        ids.append(first_sample_id(step, world, rank, B))
        keys, grads = torch.randn(B, d), torch.randn(n_live)
        stats = torch.tensor([float(rank), float(step), 0.0, 0.0])
        x = StepExchange(B, d, n_live, world, "cpu")
        x.pack(keys, grads, stats)
        g = x.all_gather()
        assert torch.equal(x.keys_of(rank), keys) and torch.equal(x.grads_of(rank), grads)
        assert [float(x.stats_of(r)[0]) for r in range(world)] == [0.0, 1.0]
        total = torch.zeros(n_live)
        for r in range(world):                     # gccb_sum_ranks: fixed rank order
            total += x.grads_of(r)
        params -= 0.1 * total / world
        for r in range(world):                     # gccb_moco_enqueue per rank, in rank order
            rows = (index + torch.arange(B)) % K
            queue[rows] = x.keys_of(r)
            index = (index + B) % K
    out[rank] = (ids, params.numpy().copy(), queue.numpy().copy(), index)
    dist.destroy_process_group()


def test_two_rank_step_exchange_keeps_replicas_identical():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    ids0, p0, q0, i0 = out[0]
    ids1, p1, q1, i1 = out[1]
    assert np.array_equal(p0, p1) and np.array_equal(q0, q1) and i0 == i1 == (3 * 2 * 4) % 32
    # sample ids: disjoint, contiguous per step across ranks -> same global stream as one big rank
    assert ids0 == [0, 8, 16] and ids1 == [4, 12, 20]
    assert np.abs(q0).sum() > 0
