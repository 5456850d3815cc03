"""Oracle (test infrastructure): Laplacian positional features of an ego-subgraph.

Follows gcc/datasets/data_util.py:242-281:
  L = D^-1/2 A D^-1/2 with D = in_degree.clip(1)      (:273-277)
  k = min(n - 2, hidden_size)                          (:278)
  top-k eigenvectors, which="LA", float64              (:245,251)
  row-L2 normalise (sklearn normalize), float32, right-pad to hidden_size (:260-262)
  k <= 0 -> zeros                                       (:243-244)

``posenc_exact`` uses a dense float64 ``eigh`` (the exact answer ARPACK
approximates).  ``posenc_reference_call`` is the reference's own scipy call
(``eigsh(..., which="LA", ncv=..., v0=rand)`` with <=10 retries doubling ncv),
used for golden cross-checks and as the CPU baseline's positional-feature stage.

The reference's output is NOT deterministic (random v0, data_util.py:248) and
its top-k cut can fall inside a degenerate eigenspace (SURVEY.md finding 5), so
parity is spectral: eigenvalues, residuals, orthonormality, and -- on columns
separated by a spectral gap -- the vectors up to sign.
"""
import numpy as np
import scipy.sparse as sparse
from scipy.sparse import linalg


def normalized_adjacency(indptr, indices, n):
    """Sparse L of data_util.py:273-277 for a sub-CSR with local ids."""
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    adj = sparse.csr_matrix((np.ones(len(indices), dtype=np.float64), indices, indptr),
                            shape=(n, n))
    in_deg = np.asarray(adj.sum(axis=0)).ravel()      # in-degree = column sums
    norm = sparse.diags(np.clip(in_deg, 1, None) ** -0.5, dtype=float)
    return (norm @ adj @ norm).tocsr()


def _finish(u, n, k, hidden_size):
    """row-L2 normalise (zero rows stay zero), f32, pad (data_util.py:260-262)."""
    nrm = np.sqrt((u * u).sum(axis=1, keepdims=True))
    nrm[nrm == 0.0] = 1.0
    x = (u / nrm).astype(np.float32)
    out = np.zeros((n, hidden_size), dtype=np.float32)
    out[:, :k] = x
    return out


def eig_topk_exact(lap, k):
    """(eigenvalues ascending [k], eigenvectors [n,k]) -- dense float64 eigh."""
    w, v = np.linalg.eigh(lap.toarray() if sparse.issparse(lap) else np.asarray(lap))
    return w[-k:], v[:, -k:]


def posenc_exact(indptr, indices, n, hidden_size=32):
    k = min(n - 2, hidden_size)
    if k <= 0:
        return np.zeros((n, hidden_size), dtype=np.float32)
    lap = normalized_adjacency(indptr, indices, n)
    _, u = eig_topk_exact(lap, k)
    return _finish(u, n, k, hidden_size)


def posenc_reference_call(indptr, indices, n, hidden_size=32, retry=10, rng=None):
    """The reference's scipy call, restated (data_util.py:242-263)."""
    k = min(n - 2, hidden_size)
    if k <= 0:
        return np.zeros((n, hidden_size), dtype=np.float32)
    lap = normalized_adjacency(indptr, indices, n).astype("float64")
    ncv = min(n, max(2 * k + 1, 20))
    rng = rng or np.random
    v0 = rng.rand(n).astype("float64") if hasattr(rng, "rand") else rng.random(n)
    u = np.zeros((n, k))
    for i in range(retry):
        try:
            _, u = linalg.eigsh(lap, k=k, which="LA", ncv=ncv, v0=v0)
        except linalg.ArpackError:
            ncv = min(ncv * 2, n)
            if i + 1 == retry:
                u = np.zeros((n, k))
        else:
            break
    return _finish(u, n, k, hidden_size)


# --- spectral parity helpers (used by tests on any candidate output) --------- #
def spectral_report(lap_dense, u):
    """For orthonormal-ish columns u [n,k]: Rayleigh quotients, residual norms,
    orthonormality defect."""
    au = lap_dense @ u
    theta = (u * au).sum(axis=0) / np.maximum((u * u).sum(axis=0), 1e-300)
    resid = np.linalg.norm(au - u * theta, axis=0)
    gram = u.T @ u
    ortho = np.abs(gram - np.eye(u.shape[1])).max()
    return theta, resid, ortho
