"""Regenerate tests/golden/egonet_cluster15.npz: the sub-CSR of one ego-net sampled by the oracle (oracle/rwr.py,
"RWR-Philox v1") from chung_lu(100000, 2000000, seed 0) with key 42, batch 32, rw_hops 256, restart 0.8 -- the one
with 173 vertices, whose normalised adjacency has the eigenvalue 1/sqrt 2 fifteen times.  Regression input of the
dense eigensolver (tests/test_emu_posenc.py).  Needs nothing but this repo."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gcc_b200.datasets import synthetic  # noqa: E402
from oracle import posenc as opos  # noqa: E402
from oracle import rwr as orwr  # noqa: E402


def main():
    g = synthetic.chung_lu(100000, 2000000, seed=0)
    B = 32
    seeds = orwr.draw_seeds(orwr.seed_cdf(g.indptr), 42, range(B))
    bt = orwr.budget_table(int(np.diff(g.indptr).max()), 256, 0.8)
    subs = orwr.rwr_batch(g.indptr, g.indices, 42, np.arange(B), seeds, bt, orwr.restart_threshold(0.8),
                          int(bt.max()) + 65, 1 << 20)
    for s in subs:
        if s["n"] != 173:
            continue
        w = np.linalg.eigvalsh(opos.normalized_adjacency(s["indptr"], s["indices"], s["n"]).toarray())
        if np.sum(np.abs(w - 2 ** -0.5) < 1e-9) == 15:
            np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "egonet_cluster15.npz"),
                     indptr=s["indptr"], indices=s["indices"])
            return
    raise SystemExit("ego-net not found")


if __name__ == "__main__":
    main()
