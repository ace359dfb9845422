"""Training meters with the interface of the reference's utils/metrics_utils.py (AverageMeter :32-59, accuracy :12-29),
as run_phase uses them (main-avid.py:144-147,174,181)."""
import collections

import torch


@torch.no_grad()
def accuracy(output, target, topk=(1,)):
    """Top-k accuracies in percent, one 1-element tensor per requested k (metrics_utils.py:12-29)."""
    ranked = output.topk(max(topk), dim=1, largest=True, sorted=True).indices          # (B, kmax) class ids, best first
    hits = ranked.eq(target.reshape(-1, 1))                                            # hit[b, j]: the j-th guess of sample b is right
    scale = 100.0 / target.size(0)
    return [hits[:, :k].any(dim=1).float().sum().reshape(1) * scale for k in topk]


class AverageMeter(object):
    """Last value and weighted mean of a scalar.  window_size > 0 averages over the last `window_size` updates only
    (metrics_utils.py:39-54); name / fmt drive __str__ exactly like the reference: 'Loss 1.234e+00 (1.111e+00)'."""

    def __init__(self, name, fmt=':f', window_size=0):
        self.name, self.fmt, self.window_size = name, fmt, window_size
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0
        self._recent = collections.deque(maxlen=self.window_size) if self.window_size > 0 else None

    @property
    def q(self):                      # the reference exposes its deque under this name
        return self._recent

    def update(self, val, n=1):
        self.val = val
        if self._recent is None:
            self.sum, self.count = self.sum + val * n, self.count + n
        else:
            self._recent.append((val, n))
            self.sum = sum(v * w for v, w in self._recent)
            self.count = sum(w for _, w in self._recent)
        self.avg = self.sum / self.count

    def __str__(self):
        template = '{name} {val%s} ({avg%s})' % (self.fmt, self.fmt)
        return template.format(name=self.name, val=self.val, avg=self.avg)
