"""Data-parallel plumbing of the pretraining step.

The reference is single-process (no collectives at all).  Here every rank holds the full CSR and
its own seeds; the ONE collective per step is an all-gather of
    [ keys k_local (B*d) | flat gradient (n_live) | stats (4) ]
after which each rank adds the gradients in fixed rank order (bit-identical replicas) and enqueues
all world*B keys in rank order (identical queues).  This module holds the host-side protocol so it
can be exercised on CPU with the gloo backend; the arithmetic on the gathered buffer runs in
libgccb200 (gccb_sum_ranks, gccb_moco_enqueue) inside PretrainEngine.
"""
import torch
import torch.distributed as dist


def first_sample_id(global_step, world_size, rank, batch):
    """Global Philox sample id of this rank's first pair at `global_step`: results are a pure
    function of (seed, sample id), hence independent of the world size."""
    return (global_step * world_size + rank) * batch


class StepExchange:
    def __init__(self, batch, dim, n_live, world_size, device, group=None):
        self.B, self.d, self.n_live, self.world, self.group = batch, dim, n_live, world_size, group
        self.payload = batch * dim + n_live + 4
        self.send = torch.zeros(self.payload, dtype=torch.float32, device=device)
        self.gathered = torch.zeros(world_size, self.payload, dtype=torch.float32, device=device)
        kd = batch * dim
        # the producers write in place: views of the send buffer
        self.keys_send = self.send[:kd].view(batch, dim)
        self.grads_send = self.send[kd:kd + n_live]
        self.stats_send = self.send[kd + n_live:]

    def pack(self, keys, grads, stats):
        """Copy into the send buffer (only needed when the producers did not write the views above)."""
        kd = self.B * self.d
        if keys.data_ptr() != self.keys_send.data_ptr():
            self.send[:kd].copy_(keys.reshape(-1))
        if grads.data_ptr() != self.grads_send.data_ptr():
            self.send[kd:kd + self.n_live].copy_(grads)
        if stats.data_ptr() != self.stats_send.data_ptr():
            self.send[kd + self.n_live:].copy_(stats)

    def all_gather(self):
        if self.send.is_cuda:
            dist.all_gather_into_tensor(self.gathered.view(-1), self.send, group=self.group)
        else:   # gloo has no all_gather_into_tensor for every torch build: use the list form
            parts = list(self.gathered.unbind(0))
            dist.all_gather(parts, self.send, group=self.group)
        return self.gathered

    # views into the gathered buffer
    def keys_of(self, rank):
        return self.gathered[rank, :self.B * self.d].view(self.B, self.d)

    def grads_of(self, rank):
        kd = self.B * self.d
        return self.gathered[rank, kd:kd + self.n_live]

    def stats_of(self, rank):
        return self.gathered[rank, self.B * self.d + self.n_live:]

    def grad_offset_bytes(self):
        return 4 * self.B * self.d
