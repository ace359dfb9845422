"""R(2+1)D video encoder on the CUDA kernels (reference: models/video.py)."""
import torch
import torch.nn as nn

from .. import ops
from ..ops import Act
from .network_blocks import flush_batch_counters, BasicR2P1DBlock, ConvBNReLU, StemOp, pad_channels
from ._tower import TowerFunction, TowerMixin


class R2Plus1D(TowerMixin, nn.Module):
    """Full 3x7x7 Conv3d stem + BN + ReLU + MaxPool3d(1,3,3), four stages of BasicR2P1DBlock, global max pool
    (video.py:12-54).  forward(x, return_embs=False) takes (B, 3, T, H, W) fp32 like the reference."""

    def __init__(self, depth=18):
        super().__init__()
        self.conv1 = nn.Sequential(
            nn.Conv3d(3, 64, kernel_size=(3, 7, 7), padding=(1, 3, 3), stride=(1, 2, 2), bias=False),
            nn.BatchNorm3d(64),
            nn.ReLU(inplace=True),
            nn.MaxPool3d(kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1)),
        )
        B = BasicR2P1DBlock
        if depth == 10:
            self.conv2x = B(64, 64)
            self.conv3x = B(64, 128, stride=(2, 2, 2))
            self.conv4x = B(128, 256, stride=(2, 2, 2))
            self.conv5x = B(256, 512, stride=(2, 2, 2))
        elif depth == 18:
            self.conv2x = nn.Sequential(B(64, 64), B(64, 64))
            self.conv3x = nn.Sequential(B(64, 128, stride=(2, 2, 2)), B(128, 128))
            self.conv4x = nn.Sequential(B(128, 256, stride=(2, 2, 2)), B(256, 256))
            self.conv5x = nn.Sequential(B(256, 512, stride=(2, 2, 2)), B(512, 512))
        elif depth == 34:
            self.conv2x = nn.Sequential(B(64, 64), B(64, 64), B(64, 64))
            self.conv3x = nn.Sequential(B(64, 128, stride=(2, 2, 2)), B(128, 128), B(128, 128), B(128, 128))
            self.conv4x = nn.Sequential(B(128, 256, stride=(2, 2, 2)), B(256, 256), B(256, 256), B(256, 256), B(256, 256), B(256, 256))
            self.conv5x = nn.Sequential(B(256, 512, stride=(2, 2, 2)), B(512, 512), B(512, 512))
        else:
            raise ValueError('depth must be 10, 18 or 34')
        self.pool = nn.AdaptiveMaxPool3d((1, 1, 1))
        self.out_dim = 512

    def _stages(self):
        for name in ('conv2x', 'conv3x', 'conv4x', 'conv5x'):
            st = getattr(self, name)
            yield name, (list(st) if isinstance(st, nn.Sequential) else [st])

    def _fwd(self, x, training, math, taps=None):
        """x (B,3,T,H,W) -> pooled (B,512); saved record for _bwd.  `taps` collects channels-last stage outputs."""
        # stem: conv -> BN -> ReLU -> max pool, the last three in one kernel (the ReLU output is never stored)
        if math == ops.MATH_FP32:
            xc = Act(ops.nchw_to_nhwc(x, c_pad=pad_channels(x.shape[1])))
            h, s_stem = ConvBNReLU.forward(xc, self.conv1[0], self.conv1[1], training, ops.MATH_FP32, pool=True)   # CUDA-core kernel
        else:   # Cin = 3: the Toeplitz-view tcgen05 stem kernel
            op = StemOp(self.conv1[0], x.shape, math)
            h, s_stem = ConvBNReLU.forward(op.pack(x), self.conv1[0], self.conv1[1], training, math, out_f32=True, op=op, pool=True)
        if taps is not None:
            taps['conv1'] = h.f32
        saved_blocks = []
        for name, blocks in self._stages():
            for blk in blocks:
                h, sb = blk._fwd(h, training, math)
                saved_blocks.append((blk, sb))
            if taps is not None:
                taps[name] = h.f32
        pooled, argmax = ops.global_maxpool_forward(h.f32)
        flush_batch_counters()
        return pooled, (s_stem, saved_blocks, argmax, tuple(h.shape))

    def _bwd(self, dpooled, saved, grads, math):
        s_stem, saved_blocks, argmax, hshape = saved
        d = ops.global_maxpool_backward(dpooled, argmax, hshape)
        sums = None
        for i in range(len(saved_blocks) - 1, -1, -1):
            blk, sb = saved_blocks[i]
            # the layer below block i: the last conv-BN-ReLU of block i-1, or the stem
            below = saved_blocks[i - 1][0].top_record(saved_blocks[i - 1][1]) if i > 0 else s_stem
            d, sums = blk._bwd(d, sb, grads, sums=sums, below=below)
        ConvBNReLU.backward(d, s_stem, grads, need_dx=False)
