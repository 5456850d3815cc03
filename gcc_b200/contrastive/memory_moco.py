"""MemoryMoCo with the reference's constructor, buffers and forward contract
(gcc/contrastive/memory_moco.py:7-63), on libgccb200 kernels (csrc/moco.cu).

Differences that do not change results: no per-step queue clone (:36), no host sync
for Z (:30, unused in the softmax branch), no host->device arange (:56).  Only the
use_softmax=True branch (what train.py:628 constructs) is implemented.
"""
import math

import torch
from torch import nn

from .. import _lib


class _MocoLogitsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, memory, T):
        lib = _lib.get()
        B, d = q.shape
        K = memory.shape[0]
        q_, k_ = q.contiguous().float(), k.contiguous().float()
        out = torch.empty(B, K + 1, dtype=torch.float32, device=q.device)
        _lib.check(lib.gccb_moco_logits(_lib.dptr(q_), _lib.dptr(k_), _lib.dptr(memory), B, d, K, T,
                                        _lib.dptr(out), _lib.stream_ptr()), "gccb_moco_logits")
        ctx.save_for_backward(k_, memory)      # caller passes a snapshot when a backward follows
        ctx.T = T
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.get()
        k_, memory = ctx.saved_tensors
        B, d = k_.shape
        K = memory.shape[0]
        dq = torch.empty(B, d, dtype=torch.float32, device=k_.device)
        _lib.check(lib.gccb_moco_logits_backward(_lib.dptr(dout.contiguous()), _lib.dptr(k_),
                                                 _lib.dptr(memory), B, d, K, ctx.T, _lib.dptr(dq),
                                                 _lib.stream_ptr()), "gccb_moco_logits_backward")
        return dq, None, None, None


class MemoryMoCo(nn.Module):
    """Fixed-size queue with momentum encoder"""

    def __init__(self, inputSize, outputSize, K, T=0.07, use_softmax=False):
        super(MemoryMoCo, self).__init__()
        if not use_softmax:
            raise NotImplementedError("only use_softmax=True (train.py:628) is implemented; the exp/Z "
                                      "branch (memory_moco.py:45-52) is dead code for pretraining")
        self.outputSize = outputSize
        self.inputSize = inputSize
        self.queueSize = K
        self.T = T
        self.index = 0
        self.use_softmax = use_softmax
        self.register_buffer("params", torch.tensor([-1]))
        stdv = 1.0 / math.sqrt(inputSize / 3)
        # same RNG call as the reference (:23) -> same queue for a given torch seed
        self.register_buffer("memory", torch.rand(self.queueSize, inputSize).mul_(2 * stdv).add_(-stdv))
        self._index_dev = None
        print("using queue shape: ({},{})".format(self.queueSize, inputSize))

    def _sync_index(self):
        dev = self.memory.device
        if self._index_dev is None or self._index_dev.device != dev:
            self._index_dev = torch.zeros(1, dtype=torch.int64, device=dev)
            self._index_dev.fill_(self.index)
        return self._index_dev

    def forward(self, q, k):
        _lib.require_device()
        lib = _lib.get()
        batchSize = q.shape[0]
        k = k.detach()
        # the logits use the queue BEFORE this batch is enqueued (memory_moco.py:36-38): the backward
        # needs that same queue, so it is snapshotted only when a backward will follow
        mem_for_logits = self.memory.clone() if (torch.is_grad_enabled() and q.requires_grad) else self.memory
        out = _MocoLogitsFn.apply(q, k, mem_for_logits, float(self.T))
        with torch.no_grad():
            idx = self._sync_index()
            kk = k.contiguous().float()
            _lib.check(lib.gccb_moco_enqueue(_lib.dptr(self.memory), _lib.dptr(kk), batchSize,
                                             self.inputSize, self.queueSize, _lib.dptr(idx), 1, 0, None, 0,
                                             _lib.stream_ptr()), "gccb_moco_enqueue")
            self.index = (self.index + batchSize) % self.queueSize
        return out
