"""NCESoftmaxLoss / NCESoftmaxLossNS (gcc/contrastive/criterions.py:5-33) on libgccb200
(csrc/moco.cu: nce_loss_kernel): mean cross-entropy of the logits against label 0 (MoCo) or
arange(B) (E2E).  Device-agnostic label construction (the reference hard-codes .cuda())."""
import torch
from torch import nn

from .. import _lib


class _NceLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, label_mode):
        lib = _lib.get()
        x_ = x.contiguous().float()
        B, Cn = x_.shape
        loss = torch.empty(1, dtype=torch.float32, device=x.device)
        dout = torch.empty_like(x_)
        _lib.check(lib.gccb_nce_loss(_lib.dptr(x_), B, Cn, label_mode, _lib.dptr(loss), _lib.dptr(dout),
                                     _lib.stream_ptr()), "gccb_nce_loss")
        ctx.save_for_backward(dout)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (dout,) = ctx.saved_tensors
        return dout * g, None


class NCESoftmaxLoss(nn.Module):
    """Softmax cross-entropy loss (a.k.a., info-NCE loss in CPC paper)"""

    def forward(self, x):
        _lib.require_device()
        x = x.squeeze()
        return _NceLossFn.apply(x, 0)


class NCESoftmaxLossNS(nn.Module):
    """Softmax cross-entropy loss (a.k.a., info-NCE loss in CPC paper), positives on the diagonal"""

    def forward(self, x):
        _lib.require_device()
        x = x.squeeze()
        return _NceLossFn.apply(x, 1)
