"""autograd plumbing shared by the two encoders: one autograd.Function per tower whose forward/backward
are straight-line sequences of libavid_b200 kernel launches (no per-op autograd nodes)."""
import os

import torch

from .. import ops

_MATH = {'fp32': ops.MATH_FP32, 'bf16x3': ops.MATH_BF16X3, 'bf16': ops.MATH_BF16}


def default_math():
    return _MATH[os.environ.get('AVID_MATH', 'bf16x3')]


def prepare_tower_filters(tower):
    """All tensor-core filters of the tower -> bf16 operand planes in one launch (instead of one launch per layer)."""
    if tower.math == ops.MATH_FP32:
        return
    ws = [m.weight.detach() for m in tower.modules()
          if isinstance(m, (torch.nn.Conv2d, torch.nn.Conv3d)) and m.in_channels % 64 == 0 and m.out_channels % 64 == 0]
    ops.prepare_filter_planes(ws, need_lo=tower.math == ops.MATH_BF16X3)


class TowerFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tower, x, *params):
        arena = tower.__dict__.setdefault('_fwd_arena', ops.ZeroArena())
        arena.begin(x.device)
        ops.set_arena(arena)
        try:
            prepare_tower_filters(tower)
            pooled, saved = tower._fwd(x.detach().contiguous().float(), tower.training, tower.math)
        finally:
            ops.set_arena(None)
            ops.clear_filter_planes()
        ctx.tower, ctx.saved, ctx.params = tower, saved, params
        return pooled

    @staticmethod
    def backward(ctx, dpooled):
        grads = {}
        arena = ctx.tower.__dict__.setdefault('_bwd_arena', ops.ZeroArena())
        arena.begin(dpooled.device)
        ops.set_arena(arena)
        ops.defer_filter_gradients(True)          # every filter gradient is brought to the PyTorch layout by ONE launch at the end
        try:
            ctx.tower._bwd(dpooled.contiguous(), ctx.saved, grads, ctx.tower.math)
            ops.flush_filter_gradients()
        finally:
            ops.defer_filter_gradients(False)
            ops.set_arena(None)
        ctx.saved = None
        return (None, None) + tuple(grads.get(p) for p in ctx.params)


class TowerMixin:
    """forward(x, return_embs=False) of models/video.py:44-54 / models/audio.py:34-44."""
    math = None

    def forward(self, x, return_embs=False):
        if self.math is None:
            self.math = default_math()
        if not x.is_cuda:
            raise RuntimeError('avid_cma_b200 encoders run on CUDA tensors only (no CPU fallback)')
        if getattr(self, '_is_replica', False):
            # nn.DataParallel over several GPUs runs its replicas in threads of ONE process; the launch bookkeeping of this package
            # (zero arena, deferred layout conversions) is per process by design: one process per GPU
            raise RuntimeError('avid_cma_b200 runs one process per GPU: use --multiprocessing-distributed / torchrun (DistributedDataParallel) '
                               'instead of multi-GPU nn.DataParallel')
        if return_embs:
            # feature taps for evaluation (utils/eval_utils.py:208,324,343): forward only, reference NC(D)HW layout
            if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and self.training:
                raise RuntimeError('return_embs=True is supported for inference (model.eval() / torch.no_grad()) only')
            taps = {}
            with torch.no_grad():
                pooled, _ = self._fwd(x.contiguous().float(), self.training, self.math, taps=taps)
            two_d = x.dim() == 4
            out = {}
            for k, v in taps.items():
                v = ops.nhwc_to_nchw(v.contiguous())
                out[k] = v.squeeze(2) if two_d else v
            out['pool'] = pooled.view(pooled.shape + ((1, 1) if two_d else (1, 1, 1)))
            return out
        params = tuple(p for p in self.parameters())
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            pooled = TowerFunction.apply(self, x, *params)
        else:
            pooled, _ = self._fwd(x.contiguous().float(), self.training, self.math)
        return pooled.view(pooled.shape + ((1, 1) if x.dim() == 4 else (1, 1, 1)))
