#!/usr/bin/env python
"""Torch-free harness for compute-sanitizer (racecheck / memcheck / synccheck) on the eigensolver kernels:
loads libgccb200.so and libcudart through ctypes, runs gccb_posenc on a handful of explicit ego-nets (every
dense class, degenerate spectra, one ChFSI ego-net) and checks the spectral bars.  Uses the test helpers of
tests/emu_util.py for the batch layout only (host arrays mirrored to the device).

    compute-sanitizer --tool racecheck python profiles/dense_sanitize.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gcc_b200 import _capi  # noqa: E402
from gcc_b200.datasets import synthetic  # noqa: E402
from oracle import posenc as opos  # noqa: E402  (checker only)

os.environ["GCCB200_DENSE_MAX"] = "228"                   # every dense class (the product default is n <= 96)
rt = C.CDLL("/usr/local/cuda/lib64/libcudart.so")
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
rt.cudaMemset.argtypes = [C.c_void_p, C.c_int, C.c_size_t]
lib = _capi.bind(C.CDLL(os.path.join(ROOT, "gcc_b200", "libgccb200.so")))


def dev(a):
    p = C.c_void_p()
    assert rt.cudaMalloc(C.byref(p), max(a.nbytes, 8)) == 0
    assert rt.cudaMemcpy(p, a.ctypes.data_as(C.c_void_p), a.nbytes, 1) == 0
    return p


def host(p, like):
    out = np.empty_like(like)
    assert rt.cudaMemcpy(out.ctypes.data_as(C.c_void_p), p, out.nbytes, 2) == 0
    return out


def main():
    import emu_util                                        # NpBatch: the host-side layout of gccb_batch_t
    small = len(sys.argv) > 1 and sys.argv[1] == "small"
    graphs = [synthetic.path_graph(5), synthetic.star_graph(33), synthetic.chung_lu(100, 250, seed=1),
              synthetic.erdos_renyi(34, 60, seed=2)]
    if not small:
        graphs += [synthetic.chung_lu(150, 400, seed=2), synthetic.star_graph(200), synthetic.erdos_renyi(228, 2000, seed=3),
                   synthetic.chung_lu(320, 900, exponent=0.7, seed=4)]
    subs = [dict(subv=np.arange(g.num_nodes, dtype=np.int32), indptr=g.indptr.astype(np.int32),
                 indices=g.indices.astype(np.int32), n=g.num_nodes, m=len(g.indices)) for g in graphs]
    half = len(subs) // 2
    views = [subs[:half], subs[half:]]
    emu_util.lib = lambda: None
    nb = emu_util.NpBatch.from_subgraphs(views)
    d = {k: dev(getattr(nb, k)) for k in ("node_off", "edge_off", "indptr", "indices", "sub_deg", "graph_id", "orig_id",
                                          "counters", "flags")}
    batch = _capi.Batch(nb.B, nb.node_cap, nb.edge_cap, 0, d["node_off"], d["edge_off"], d["indptr"], d["indices"],
                        d["sub_deg"], d["graph_id"], d["orig_id"], d["counters"], d["flags"])
    pos = np.zeros((2, nb.node_cap, 32), np.float32)
    eig = np.zeros((2 * nb.B, 32), np.float32)
    ws = np.zeros(lib.gccb_posenc_workspace(nb.B, nb.node_cap), np.uint8)
    dpos, deig, dws = dev(pos), dev(eig), dev(ws)
    rc = lib.gccb_posenc(C.byref(batch), 32, 0, dpos, deig, dws, ws.nbytes, None)
    assert rc == 0, lib.gccb_last_error()
    assert rt.cudaDeviceSynchronize() == 0
    pos, eig = host(dpos, pos), host(deig, eig)
    flags = host(d["flags"], nb.flags)
    worst = 0.0
    for v in (0, 1):
        for gi, s in enumerate(views[v]):
            a, z = nb.node_off[v, gi], nb.node_off[v, gi + 1]
            n = s["n"]
            k = min(n - 2, 32)
            lap = opos.normalized_adjacency(s["indptr"], s["indices"], n).toarray()
            w, _ = opos.eig_topk_exact(lap, k)
            theta, resid, ortho = opos.spectral_report(lap, pos[v, a:z, :k].astype(np.float64))
            bar = 2e-5 if n <= 228 else 2.5e-3
            assert np.abs(eig[v * nb.B + gi, :k] - w).max() < (2e-6 if n <= 228 else 1e-3) and resid.max() < bar and ortho < 1e-4, (
                n, np.abs(eig[v * nb.B + gi, :k] - w).max(), resid.max(), ortho)
            worst = max(worst, resid.max())
    print("dense_sanitize ok: %d ego-nets, flags %d, worst residual %.2e" % (len(subs), int(flags[0]), worst))


if __name__ == "__main__":
    main()
