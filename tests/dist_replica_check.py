"""Launched by torchrun on 2 GPUs (tests/test_gpu_parity2.py): data-parallel replica identity.
Each rank trains 5 steps; then (1) parameters / EMA parameters / queue / queue pointer are bit-identical
across ranks, (2) the first step's summed gradient equals, bit for bit, the fixed-order sum of the gradients
two single-GPU engines produce on the two shards (sample ids (step*2 + rank)*B + i)."""
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build(rank, world, pg_on):
    from gcc_b200.contrastive.memory_moco import MemoryMoCo
    from gcc_b200.datasets import synthetic
    from gcc_b200.datasets.graph_dataset import LoadBalanceGraphDataset
    from gcc_b200.engine import PretrainEngine
    from gcc_b200.models import GraphEncoder
    torch.manual_seed(0)
    g = synthetic.chung_lu(20000, 200000, seed=1)
    B, H, L, K = 32, 64, 3, 256
    ds = LoadBalanceGraphDataset(rw_hops=64, restart_prob=0.8, dgl_graphs_file=g, num_samples=256, batch_size=B, seed=7)

    def mk():
        return GraphEncoder(positional_embedding_size=32, max_degree=512, degree_embedding_size=16, output_dim=H,
                            node_hidden_dim=H, num_layers=L, norm=True, gnn_model="gin", degree_input=True)

    model, ema = mk(), mk()
    ema.load_state_dict(model.state_dict())
    model, ema = model.cuda(), ema.cuda()
    contrast = MemoryMoCo(H, None, K, 0.07, use_softmax=True).cuda()
    eng = PretrainEngine(ds, model, ema, contrast, moco=True, rank=rank, world_size=world if pg_on else 1, prefetch=2)
    return eng, model, ema, contrast


def digest(*tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(t.detach().cpu().numpy().tobytes())
    return h.hexdigest()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    eng, model, ema, contrast = build(rank, world, True)
    eng.step(lr=0.005)
    torch.cuda.synchronize()
    g_sum = eng.grads.detach().clone()                      # after gccb_sum_ranks: sum over ranks, fixed order
    # the two shards on ONE process each, no collective: sample ids must be the ones the ranks used
    from gcc_b200.parallel import first_sample_id
    parts = []
    for r in range(world):
        e1, m1, _, _ = build(0, 1, False)
        first = first_sample_id(0, world, r, e1.B)
        buf = e1.ds.sample_batch(first_sample=first)
        e1.step(lr=0.005, _presampled=True)
        torch.cuda.synchronize()
        parts.append(e1.grads.detach().clone())
    want = parts[0] + parts[1]                              # rank order 0, 1 (gccb_sum_ranks)
    assert torch.allclose(g_sum, want, rtol=1e-4, atol=1e-6), float((g_sum - want).abs().max())   # (BatchNorm sums are float64 atomics: last-bit freedom)
    for _ in range(4):
        eng.step(lr=0.005)
    eng.read_stats()
    d = digest(model.flat_params, ema.flat_params, contrast.memory, eng.index_dev, eng.adam_m, eng.adam_v)
    all_d = [None] * world
    dist.all_gather_object(all_d, d)
    assert len(set(all_d)) == 1, all_d
    if rank == 0:
        print("REPLICAS IDENTICAL", d[:16], "index", int(eng.index_dev.item()))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
