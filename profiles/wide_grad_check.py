#!/usr/bin/env python
"""GPU diagnostic: per-parameter gradient error of the wide (hidden 256 / 128) GIN encoder against the fp64
oracle, module API, for the fp32 SIMT path and the tensor-core path; run twice to expose nondeterminism."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gcc_b200.datasets import synthetic  # noqa: E402
from gcc_b200.datasets.data_util import BatchedSubgraphs  # noqa: E402
from gcc_b200.models import GraphEncoder  # noqa: E402
from oracle import model as om  # noqa: E402
from test_gpu_parity import _dataset  # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = synthetic.chung_lu(4000, 30000, seed=6)
B, L = 24, 5
ds = _dataset(g, B, 64, seed=3)
buf = ds.sample_batch(first_sample=0)
torch.cuda.synchronize()
n, m = int(buf.node_off[0, B]), int(buf.edge_off[0, B])
noff = buf.node_off[0].cpu().numpy().astype(np.int64)
seed = np.zeros(n, np.int64)
seed[noff[:B]] = 1
args = (buf.indptr[0, :n + 1].cpu().numpy().astype(np.int64), buf.indices[0, :m].cpu().numpy().astype(np.int64),
        buf.pos[0, :n].cpu().double(), seed, buf.sub_deg[0, :n].cpu().numpy(), noff)
print("N", n, "max induced degree", int(buf.sub_deg[0, :n].max()))
for tc in (0, 1):
    prev = None
    for rep in range(2):
        torch.manual_seed(5)
        model = GraphEncoder(positional_embedding_size=32, max_degree=512, degree_embedding_size=16, output_dim=H,
                             node_hidden_dim=H, num_layers=L, norm=True, gnn_model="gin", degree_input=True)
        model.cfg.tensor_cores = tc
        model = model.cuda()
        model.train()
        model.gnn.drop.eval()
        sd0 = {k: v.detach().cpu().double().clone() for k, v in model.state_dict().items()}
        feat = model(BatchedSubgraphs(buf, 0))
        R = torch.randn(B, H, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
        (feat * R).sum().backward()
        torch.cuda.synchronize()
        grads = {k: p.grad.detach().cpu().numpy().copy() for k, p in model.named_parameters() if p.grad is not None}
        if prev is not None:
            print("tc=%d run-to-run max |diff| over all grads: %.3e" % (tc, max(np.abs(grads[k] - prev[k]).max() for k in grads)))
        prev = grads
    P = {k: (v.clone().requires_grad_(True) if not k.endswith(("running_mean", "running_var", ".eps", "num_batches_tracked")) else v)
         for k, v in sd0.items()}
    f, _, _ = om.gin_encoder_forward(P, *args, num_layers=L, bn_train=True, gemm_operand_dtype=torch.bfloat16 if tc else None)
    (f * R.cpu().double()).sum().backward()
    print("tc=%d feat max|diff| %.2e" % (tc, np.abs(feat.detach().cpu().numpy() - f.detach().numpy()).max()))
    for k, gv in grads.items():
        if k.startswith(("set2set", "lin_readout")) or ("mlp.linears" in k and k.endswith("bias")):
            continue
        want = P[k].grad
        want = np.zeros(gv.shape) if want is None else want.numpy()
        scale = max(np.abs(want).max(), 1e-6)
        e = np.abs(gv - want).max() / scale
        if e > 2e-3:
            i = np.unravel_index(np.abs(gv - want).argmax(), gv.shape)
            print("  tc=%d %-52s err/scale %.2e scale %.2e at %s got %.4f want %.4f" % (tc, k, e, scale, i, gv[i], want[i]))
