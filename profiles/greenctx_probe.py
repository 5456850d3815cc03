#!/usr/bin/env python
"""GPU feasibility probe: SM partitioning with CUDA green contexts + torch external streams."""
import sys
import time

import torch
from cuda.bindings import driver as drv


def ck(res):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError("CUDA driver error %s" % err)
    return res[1:] if len(res) > 2 else (res[1] if len(res) == 2 else None)


torch.cuda.init()
x = torch.zeros(1, device="cuda")        # primary context is current
dev = ck(drv.cuDeviceGet(0))
res = ck(drv.cuDeviceGetDevResource(dev, drv.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
print("device SMs:", res.sm.smCount)
n_train = int(sys.argv[1]) if len(sys.argv) > 1 else 48
out = drv.cuDevSmResourceSplitByCount(1, res, 0, n_train)
print("split ->", [int(out[0])], "groups", out[2], "first", out[1][0].sm.smCount, "remaining", out[3].sm.smCount)
grp, rem = out[1][0], out[3]
streams = []
for r in (grp, rem):
    desc = ck(drv.cuDevResourceGenerateDesc([r], 1))
    g = ck(drv.cuGreenCtxCreate(desc, dev, drv.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
    s = ck(drv.cuGreenCtxStreamCreate(g, drv.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
    streams.append(torch.cuda.ExternalStream(int(s)))
    print("green ctx with", ck(drv.cuGreenCtxGetDevResource(g, drv.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM)).sm.smCount, "SMs; stream", hex(int(s)))
sa, sb = streams
a = torch.randn(8192, 8192, device="cuda")
b = torch.randn(8192, 8192, device="cuda")
torch.cuda.synchronize()


def timed(stream, n=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
        for _ in range(n):
            c = a @ b
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("matmul ms: default %.3f, green A (%d SMs) %.3f, green B %.3f" % (
    timed(torch.cuda.current_stream()), n_train, timed(sa), timed(sb)))
# cross-stream events between green streams and the default stream
ev = torch.cuda.Event()
with torch.cuda.stream(sa):
    c = a @ b
    ev.record()
sb.wait_event(ev)
with torch.cuda.stream(sb):
    d = c.sum()
torch.cuda.current_stream().wait_stream(sb)
print("cross-stream ok:", float(d) == float((a @ b).sum()) or abs(float(d) - float((a @ b).sum())) < 1e-2 * abs(float(d)))
# our library on a green stream (sampler + posenc with its internal side streams)
sys.path.insert(0, ".")
import bench
from gcc_b200.datasets.graph_dataset import LoadBalanceGraphDataset
cfg = bench.CONFIGS["c2"]
g = bench.make_graph_device(cfg, torch.device("cuda"))
ds = LoadBalanceGraphDataset(rw_hops=cfg["rw_hops"], restart_prob=0.8, dgl_graphs_file=g, batch_size=256, seed=0)
for st, nm in ((torch.cuda.current_stream(), "default"), (sb, "green B"), (sa, "green A")):
    with torch.cuda.stream(st):
        for _ in range(2):
            ds.sample_batch(first_sample=0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ds.sample_batch(first_sample=0)
        e1.record()
    torch.cuda.synchronize()
    ds.buffers.check_flags()
    print("sample+posenc on %s: %.3f ms" % (nm, e0.elapsed_time(e1)))
