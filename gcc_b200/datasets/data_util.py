"""Batched ego-net view object: the stand-in for the batched DGLGraph that
gcc/datasets/data_util.py:26-32 (batcher -> dgl.batch) hands to the model.

It exposes exactly the surface GraphEncoder.forward and train_moco touch
(SURVEY.md section 8b): .ndata["pos_undirected"], .ndata["seed"], .in_degrees(),
.batch_size, .number_of_nodes(), .number_of_edges(), .to(device).  Data stays
in the device buffers the sampler wrote; size queries are the only host syncs.
"""
import torch


class _NData(dict):
    def __init__(self, owner):
        super().__init__()
        self._owner = owner

    def __missing__(self, key):
        o = self._owner
        n = o.number_of_nodes()
        if key == "pos_undirected":
            val = o.buffers.pos[o.view, :n]
        elif key == "seed":                     # one-hot of the first row of each graph (data_util.py:234-238)
            val = torch.zeros(n, dtype=torch.long, device=o.buffers.pos.device)
            val[o.buffers.node_off[o.view, :o.batch_size].long()] = 1
        else:
            raise KeyError(key)
        self[key] = val
        return val


class BatchedSubgraphs:
    def __init__(self, buffers, view):
        self.buffers, self.view = buffers, view
        self.batch_size = buffers.B
        self.ndata = _NData(self)
        self._n = self._m = None

    def _sizes(self):
        if self._n is None:
            self.buffers.check_flags()
            self._n = int(self.buffers.node_off[self.view, self.batch_size].item())
            self._m = int(self.buffers.edge_off[self.view, self.batch_size].item())
        return self._n, self._m

    def number_of_nodes(self):
        return self._sizes()[0]

    def number_of_edges(self):
        return self._sizes()[1]

    @property
    def batch_num_nodes(self):
        off = self.buffers.node_off[self.view]
        return (off[1:] - off[:-1]).tolist()

    def in_degrees(self):
        return self.buffers.sub_deg[self.view, :self.number_of_nodes()].long()

    def to(self, device):                      # train.py:382-383 -- already resident
        return self

    def csr(self):
        n, m = self._sizes()
        return self.buffers.indptr[self.view, :n + 1], self.buffers.indices[self.view, :m]


def batcher():
    """API parity with gcc/datasets/data_util.py:26-32: the device dataset already yields
    batched pairs, so the collate function is the identity on a 1-item list."""
    def batcher_dev(batch):
        return batch[0] if isinstance(batch, list) and len(batch) == 1 else batch
    return batcher_dev
