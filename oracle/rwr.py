"""Oracle (test infrastructure): seed draw, RWR walk, ego-subgraph induction.

Two restatements of the same spec ("RWR-Philox v1"):
  * the C one in gccb_oracle.c (fast; also the CPU baseline), via ctypes;
  * a pure-Python one here (slow, small cases only) used to cross-check the C.

Reference lines followed (the glue the reference owns):
  gcc/datasets/graph_dataset.py:85-92   seed draw  p ~ in_deg^0.75
  gcc/datasets/graph_dataset.py:113-124 walk budget
  gcc/datasets/graph_dataset.py:125-130 RWR from [seed, seed] (step_dist=[1,0,0])
  gcc/datasets/data_util.py:218-239     node order, induction, seed one-hot
DGL-owned semantics are documented choices -> "parity unpinned" vs DGL.
"""
import ctypes
import math

import numpy as np

from . import build as _build

HOPCAP = 64
TAG_WALK, TAG_SEED, TAG_DROPOUT = 0, 1, 2
M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK32 = 0xFFFFFFFF

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
        _lib.gccb_o_draw_seed.restype = ctypes.c_int64
        _lib.gccb_o_draw_seed.argtypes = [ctypes.c_void_p, ctypes.c_int64,
                                          ctypes.c_uint64, ctypes.c_uint64]
        _lib.gccb_o_rwr_subgraph.restype = ctypes.c_int
        _lib.gccb_o_rwr_subgraph.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64,
            ctypes.c_uint64, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
            ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        _lib.gccb_o_rwr_batch.restype = ctypes.c_int
        _lib.gccb_o_rwr_batch.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
            ctypes.c_uint32, ctypes.c_int64, ctypes.c_int32, ctypes.c_int64,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib.gccb_o_philox4x32_10.restype = None
        _lib.gccb_o_philox4x32_10.argtypes = [ctypes.c_void_p] * 3
        _lib.gccb_o_dropout_mask.restype = None
        _lib.gccb_o_dropout_mask.argtypes = [
            ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int64,
            ctypes.c_uint32, ctypes.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------- #
# host-side tables (shared spec; the product builds its own copies)
# --------------------------------------------------------------------------- #
def restart_threshold(restart_prob):
    """floor(restart_prob * 2^32) clamped to u32."""
    return min(int(restart_prob * 4294967296.0), MASK32)


def budget_for_degree(deg, rw_hops, restart_prob):
    """graph_dataset.py:113-124, verbatim arithmetic."""
    return max(rw_hops,
               int(((deg ** 0.75) * math.e / (math.e - 1) / restart_prob) + 0.5))


def budget_table(max_deg, rw_hops, restart_prob):
    return np.array([budget_for_degree(d, rw_hops, restart_prob)
                     for d in range(max_deg + 1)], dtype=np.int32)


def seed_cdf(indptr):
    """graph_dataset.py:85-87: prob = deg^0.75 / sum; cumulative, last = 1."""
    deg = np.diff(np.asarray(indptr, dtype=np.int64)).astype(np.float64)
    p = deg ** 0.75
    p = p / p.sum()
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    return cdf


# --------------------------------------------------------------------------- #
# C oracle wrappers
# --------------------------------------------------------------------------- #
def philox_c(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().gccb_o_philox4x32_10(_p(c), _p(k), _p(o))
    return o


def draw_seeds(cdf, key, sample_ids):
    cdf = np.ascontiguousarray(cdf, dtype=np.float64)
    return np.array([lib().gccb_o_draw_seed(_p(cdf), len(cdf), key, int(s))
                     for s in sample_ids], dtype=np.int64)


def rwr_subgraph(indptr, indices, key, sample, view, seed, budget, restart_thresh,
                 cap_n=None, cap_m=None):
    """Returns dict(subv, indptr, indices, n, m, steps, traces, sumdeg)."""
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    cap_n = int(cap_n or (budget + HOPCAP + 1))
    cap_m = int(cap_m or min(cap_n * cap_n, 1 << 26))
    subv = np.zeros(cap_n, dtype=np.int32)
    sp = np.zeros(cap_n + 1, dtype=np.int32)
    si = np.zeros(cap_m, dtype=np.int32)
    cnt = np.zeros(5, dtype=np.int64)
    rc = lib().gccb_o_rwr_subgraph(_p(indptr), _p(indices), len(indptr) - 1, key,
                                   sample, view, seed, budget, restart_thresh,
                                   _p(subv), cap_n, _p(sp), _p(si), cap_m, _p(cnt))
    if rc:
        raise RuntimeError("oracle rwr_subgraph rc=%d" % rc)
    n, m = int(cnt[0]), int(cnt[1])
    return dict(subv=subv[:n].copy(), indptr=sp[:n + 1].copy(), indices=si[:m].copy(),
                n=n, m=m, steps=int(cnt[2]), traces=int(cnt[3]), sumdeg=int(cnt[4]))


def rwr_batch(indptr, indices, key, sample_ids, seeds, btable, restart_thresh,
              cap_n, cap_m):
    """B samples x 2 views; slot s = 2*i + view.  Returns list of dicts."""
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    sample_ids = np.ascontiguousarray(sample_ids, dtype=np.int64)
    seeds = np.ascontiguousarray(seeds, dtype=np.int64)
    btable = np.ascontiguousarray(btable, dtype=np.int32)
    B = len(seeds)
    subv = np.empty((2 * B, cap_n), dtype=np.int32)      # empty, not zeros: only [:n] / [:m] is read back
    sp = np.empty((2 * B, cap_n + 1), dtype=np.int32)
    si = np.empty((2 * B, cap_m), dtype=np.int32)
    cnt = np.zeros((2 * B, 5), dtype=np.int64)
    rc = lib().gccb_o_rwr_batch(_p(indptr), _p(indices), len(indptr) - 1, key,
                                _p(sample_ids), _p(seeds), _p(btable), len(btable),
                                restart_thresh, B, cap_n, cap_m, _p(subv), _p(sp),
                                _p(si), _p(cnt))
    if rc:
        raise RuntimeError("oracle rwr_batch rc=%d" % rc)
    out = []
    for s in range(2 * B):
        n, m = int(cnt[s, 0]), int(cnt[s, 1])
        out.append(dict(subv=subv[s, :n].copy(), indptr=sp[s, :n + 1].copy(),
                        indices=si[s, :m].copy(), n=n, m=m, steps=int(cnt[s, 2]),
                        traces=int(cnt[s, 3]), sumdeg=int(cnt[s, 4])))
    return out


def dropout_mask(key, step, layer, count, p=0.5):
    keep = np.zeros(count, dtype=np.uint8)
    thresh = min(int((1.0 - p) * 4294967296.0), MASK32)
    lib().gccb_o_dropout_mask(key, step, layer, count, thresh, _p(keep))
    return keep.astype(bool)


# --------------------------------------------------------------------------- #
# pure-Python restatement (independent of the C; small cases only)
# --------------------------------------------------------------------------- #
def philox_py(ctr, key):
    c0, c1, c2, c3 = [int(x) for x in ctr]
    k0, k1 = [int(x) for x in key]
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK32, p1 & MASK32, \
                         ((p0 >> 32) ^ c3 ^ k1) & MASK32, p0 & MASK32
        k0 = (k0 + W0) & MASK32
        k1 = (k1 + W1) & MASK32
    return [c0, c1, c2, c3]


def _philox_at(key, sample, trace, hop, view, tag):
    ctr = [sample & MASK32, (sample >> 32) & MASK32, trace & MASK32,
           (hop | (view << 8) | (tag << 16)) & MASK32]
    return philox_py(ctr, [key & MASK32, (key >> 32) & MASK32])


def draw_seed_py(cdf, key, sample):
    w = _philox_at(key, sample, 0, 0, 0, TAG_SEED)
    u = float((w[0] << 21) | (w[1] >> 11)) / 9007199254740992.0
    i = int(np.searchsorted(cdf, u, side="right"))
    return min(i, len(cdf) - 1)


def rwr_traces_py(indptr, indices, key, sample, view, seed, budget, restart_thresh):
    """List of traces (each a list of node ids, seed excluded) -- the shape DGL's
    random_walk_with_restart returns per seed (list of tensors)."""
    traces, total, t = [], 0, 0
    while total < budget:
        cur, tr = seed, []
        for hop in range(HOPCAP):
            w = _philox_at(key, sample, t, hop, view, TAG_WALK)
            if hop > 0 and w[0] < restart_thresh:
                break
            beg, deg = int(indptr[cur]), int(indptr[cur + 1] - indptr[cur])
            assert deg > 0, "no successors from vertex"
            cur = int(indices[beg + ((w[1] * deg) >> 32)])
            tr.append(cur)
        traces.append(tr)
        total += len(tr)
        t += 1
    return traces


def induce_py(indptr, indices, seed, traces):
    """data_util.py:221-230 with g.subgraph semantics (ids follow subv order)."""
    subv = sorted(set(v for tr in traces for v in tr))
    if seed in subv:
        subv.remove(seed)
    subv = [seed] + subv
    pos = {v: i for i, v in enumerate(subv)}
    sp, si = [0], []
    for v in subv:
        for e in range(int(indptr[v]), int(indptr[v + 1])):
            j = pos.get(int(indices[e]))
            if j is not None:
                si.append(j)
        sp.append(len(si))
    return (np.array(subv, dtype=np.int32), np.array(sp, dtype=np.int32),
            np.array(si, dtype=np.int32))


def rwr_subgraph_py(indptr, indices, key, sample, view, seed, budget, restart_thresh):
    traces = rwr_traces_py(indptr, indices, key, sample, view, seed, budget,
                           restart_thresh)
    subv, sp, si = induce_py(indptr, indices, seed, traces)
    return dict(subv=subv, indptr=sp, indices=si, n=len(subv), m=len(si),
                steps=sum(len(t) for t in traces), traces=len(traces))
