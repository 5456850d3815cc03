"""LR schedule and meters: behaviour of gcc/utils/misc.py:5-42 (host-side scalars, no kernels)."""
import numpy as np


def warmup_linear(x, warmup=0.002):
    """Triangular schedule: linear warm-up to 1 at x == warmup, then linear decay to 0 at x == 1
    (gcc/utils/misc.py:5-10)."""
    if x < warmup:
        return x / warmup
    return max((x - 1.0) / (warmup - 1.0), 0)


def adjust_learning_rate(epoch, opt, optimizer):
    """Step decay by opt.lr_decay_rate at opt.lr_decay_epochs (gcc/utils/misc.py:13-19)."""
    steps = np.sum(epoch > np.asarray(opt.lr_decay_epochs))
    if steps > 0:
        new_lr = opt.learning_rate * (opt.lr_decay_rate ** steps)
        for param_group in optimizer.param_groups:
            param_group["lr"] = new_lr


class AverageMeter(object):
    """Running value / average (gcc/utils/misc.py:22-42)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count
