"""CPU: eigensolver kernel (gcc_b200/csrc/posenc.cu) under the fiber emulator --
spectral parity with the oracle (dense float64 eigh) and with the reference's own
outputs (tests/golden/posenc_golden.npz).  Kernel LOGIC only; see test_gpu_*."""
import ctypes as C
import os

import numpy as np
import pytest

from emu_util import NpBatch, lib, ptr
from gcc_b200.datasets import synthetic
from oracle import posenc as opos

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _sub(g):
    return dict(subv=np.arange(g.num_nodes, dtype=np.int32), indptr=g.indptr.astype(np.int32),
                indices=g.indices.astype(np.int32), n=g.num_nodes, m=len(g.indices))


def _posenc(views, normalize):
    L = lib()
    b = NpBatch.from_subgraphs(views)
    pos = np.full((2, b.node_cap, 32), np.nan, np.float32)
    eig = np.full((2 * b.B, 32), np.nan, np.float32)
    ws = np.zeros(L.gccb_posenc_workspace(b.B, b.node_cap), np.uint8)
    rc = L.gccb_posenc(C.byref(b.c), 32, normalize, ptr(pos), ptr(eig), ptr(ws), ws.nbytes, None)
    assert rc == 0, L.gccb_last_error()
    return b, pos, eig


def _check_spectral(sub, u, lam):
    n = sub["n"]
    k = min(n - 2, 32)
    lap = opos.normalized_adjacency(sub["indptr"], sub["indices"], n).toarray()
    if k <= 0:
        assert np.all(u == 0) and np.all(lam == 0)
        return
    w_exact, _ = opos.eig_topk_exact(lap, k)
    assert np.all(u[:, k:] == 0) and np.all(lam[k:] == 0)
    assert np.allclose(lam[:k], w_exact, atol=1e-5), np.abs(lam[:k] - w_exact).max()
    theta, resid, ortho = opos.spectral_report(lap, u[:, :k].astype(np.float64))
    assert resid.max() < 1e-4, resid.max()
    assert ortho < 1e-4, ortho
    assert np.allclose(theta, w_exact, atol=1e-5)


def test_jacobi_spectral_parity_small_and_degenerate():
    graphs = [synthetic.path_graph(2), synthetic.path_graph(3), synthetic.path_graph(9),
              synthetic.star_graph(20), synthetic.triangle_tail(4),
              synthetic.erdos_renyi(40, 90, seed=1), synthetic.star_graph(50),
              synthetic.erdos_renyi(60, 100, seed=4)]
    half = len(graphs) // 2
    views = [[_sub(g) for g in graphs[:half]], [_sub(g) for g in graphs[half:]]]
    b, pos, eig = _posenc(views, normalize=0)
    assert b.flags[0] == 0
    for v in (0, 1):
        for gi, sub in enumerate(views[v]):
            a, z = b.node_off[v, gi], b.node_off[v, gi + 1]
            _check_spectral(sub, pos[v, a:z], eig[v * b.B + gi])


def test_jacobi_size_classes_and_normalisation():
    g1 = synthetic.erdos_renyi(90, 240, seed=7)           # 64 < n <= 96: Chebyshev-filtered subspace iteration
    g2 = synthetic.star_graph(90)                         # extreme degeneracy (eigenvalue 0 x 89)
    g3 = synthetic.erdos_renyi(150, 420, seed=9)          # 96 < n <= 160: second shared-memory class
    views = [[_sub(g1), _sub(g3)], [_sub(g2), _sub(synthetic.path_graph(30))]]
    assert 64 < g1.num_nodes <= 96 < g3.num_nodes <= 160 and 64 < g2.num_nodes
    b, pos, eig = _posenc(views, normalize=0)
    assert b.flags[0] == 0
    for v in (0, 1):
        for gi, sub in enumerate(views[v]):
            a, z = b.node_off[v, gi], b.node_off[v, gi + 1]
            _check_spectral(sub, pos[v, a:z], eig[v * b.B + gi])
    b, posn, _ = _posenc(views, normalize=1)
    for v in (0, 1):
        n = b.node_off[v, b.B]
        assert np.allclose(np.linalg.norm(posn[v, :n], axis=1), 1.0, atol=1e-5)
        raw = pos[v, :n]
        want = raw / np.linalg.norm(raw, axis=1, keepdims=True)
        assert np.allclose(posn[v, :n], want, atol=1e-6)


def test_posenc_matches_reference_golden():
    z = np.load(os.path.join(G, "posenc_golden.npz"))
    subs = []
    for ci in range(int(z["num_cases"])):
        ip, ix = z["indptr%d" % ci].astype(np.int32), z["indices%d" % ci].astype(np.int32)
        subs.append(dict(subv=np.arange(len(ip) - 1, dtype=np.int32), indptr=ip, indices=ix,
                         n=len(ip) - 1, m=len(ix)))
    if len(subs) % 2:
        subs.append(subs[0])
    half = len(subs) // 2
    views = [subs[:half], subs[half:]]
    b, pos, _ = _posenc(views, normalize=1)
    checked = 0
    for ci in range(int(z["num_cases"])):
        v, gi = (0, ci) if ci < half else (1, ci - half)
        a, zz = b.node_off[v, gi], b.node_off[v, gi + 1]
        got, want = pos[v, a:zz], z["pos%d" % ci]
        n = zz - a
        k = min(n - 2, 32)
        if k <= 0:
            assert np.all(got == 0) and np.all(want == 0)
        elif bool(z["simple%d" % ci]):
            s = np.sign((got[:, :k] * want[:, :k]).sum(axis=0))
            assert np.allclose(got[:, :k] * s, want[:, :k], atol=5e-5), (ci, np.abs(got[:, :k] * s - want[:, :k]).max())
            assert np.all(got[:, k:] == 0)
            checked += 1
    assert checked >= 8


def test_huge_egonet_one_block_in_shared_memory():
    g = synthetic.chung_lu(560, 1500, exponent=0.8, seed=3)    # 480 < n <= 1000: X in smem, Y in workspace
    assert 480 < g.num_nodes <= 1000
    views = [[_sub(g)], [_sub(synthetic.path_graph(5))]]
    b, pos, eig = _posenc(views, normalize=0)
    assert b.flags[0] == 0
    _check_spectral(views[0][0], pos[0, :g.num_nodes], eig[0])


def test_large_egonet_goes_through_chfsi():
    g = synthetic.chung_lu(260, 700, seed=3)              # hub-and-leaves: large degenerate cluster
    assert g.num_nodes > 200
    views = [[_sub(g)], [_sub(synthetic.path_graph(5))]]
    b, pos, eig = _posenc(views, normalize=0)
    assert b.flags[0] == 0
    _check_spectral(views[0][0], pos[0, :g.num_nodes], eig[0])
    _check_spectral(views[1][0], pos[1, :5], eig[1])
