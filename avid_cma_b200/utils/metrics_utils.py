"""Training meters with the interface of the reference's utils/metrics_utils.py (AverageMeter :32-59, accuracy :12-29),
as run_phase uses them (main-avid.py:144-147,174,181)."""
import collections

import torch


def accuracy(output, target, topk=(1,)):
    """Top-k accuracies in percent (metrics_utils.py:12-29): one 1-element tensor per k."""
    with torch.no_grad():
        kmax = max(topk)
        pred = output.topk(kmax, dim=1, largest=True, sorted=True).indices          # (B, kmax)
        hit = pred.eq(target.view(-1, 1))
        n = target.size(0)
        return [hit[:, :k].reshape(-1).float().sum(0, keepdim=True) * (100.0 / n) for k in topk]


class AverageMeter(object):
    """Last value and (optionally windowed) weighted mean.  window_size > 0 keeps the last `window_size` updates only
    (metrics_utils.py:39-54); name / fmt drive __str__ exactly like the reference ('Loss 1.234e+00 (1.111e+00)')."""

    def __init__(self, name, fmt=':f', window_size=0):
        self.name, self.fmt, self.window_size = name, fmt, window_size
        self.reset()

    def reset(self):
        self._window = collections.deque(maxlen=self.window_size) if self.window_size > 0 else None
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        if self._window is not None:
            self._window.append((val, n))
            self.count = sum(w for _, w in self._window)
            self.sum = sum(v * w for v, w in self._window)
        else:
            self.count += n
            self.sum += val * n
        self.avg = self.sum / self.count

    @property
    def q(self):                      # the reference exposes its deque under this name
        return self._window

    def __str__(self):
        return ('{name} {val' + self.fmt + '} ({avg' + self.fmt + '})').format(name=self.name, val=self.val, avg=self.avg)
