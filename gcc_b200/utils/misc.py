"""Host-side scalars of the training loop: the triangular LR schedule, the (unused by pretraining)
epoch-step decay and the running meter -- same names and behaviour as gcc/utils/misc.py:5-42,
pinned by tests/golden/misc_golden.npz.  No kernels here."""


def warmup_linear(x, warmup=0.002):
    """Fraction of the peak learning rate at training progress x in [0, 1]: a straight line from 0
    up to 1 at x == warmup, then a straight line down to 0 at x == 1, clamped at 0 beyond."""
    rising = x < warmup
    scale = x / warmup if rising else (1.0 - x) / (1.0 - warmup)
    return scale if rising else max(scale, 0)


def adjust_learning_rate(epoch, opt, optimizer):
    """Multiply opt.learning_rate by opt.lr_decay_rate once per milestone in opt.lr_decay_epochs that
    `epoch` has passed, and write it into every parameter group (no-op before the first one)."""
    passed = sum(1 for milestone in opt.lr_decay_epochs if epoch > milestone)
    if passed == 0:
        return
    decayed = opt.learning_rate * opt.lr_decay_rate ** passed
    for group in optimizer.param_groups:
        group["lr"] = decayed


class AverageMeter(object):
    """Last value (`val`), weighted total (`sum`), weight (`count`) and mean (`avg`) of a stream of
    numbers; train.py reads `.val` and `.avg` when it prints."""

    __slots__ = ("val", "avg", "sum", "count")

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.count += n
        self.sum += n * val
        self.avg = self.sum / self.count
