"""Conv2D audio encoder on the CUDA kernels (reference: models/audio.py)."""
import torch.nn as nn

from .. import ops
from ..ops import Act
from .network_blocks import flush_batch_counters, Basic2DBlock, ConvBNReLU, StemOp, pad_channels
from ._tower import TowerMixin

__all__ = ['Conv2D']


class Conv2D(TowerMixin, nn.Module):
    """7x7 s2 stem + four Basic2DBlocks + global max pool (audio.py:15-44); input (B, 1, T, F) fp32."""

    def __init__(self, depth=10):
        super().__init__()
        assert depth == 10
        self.conv1 = nn.Sequential(
            nn.Conv2d(1, 64, kernel_size=7, padding=3, stride=2, bias=False),
            nn.BatchNorm2d(64),
            nn.ReLU(inplace=True),
        )
        self.block1 = Basic2DBlock(64, 64, stride=(2, 2))
        self.block2 = Basic2DBlock(64, 128, stride=(2, 2))
        self.block3 = Basic2DBlock(128, 256, stride=(2, 2))
        self.block4 = Basic2DBlock(256, 512)
        self.pool = nn.AdaptiveMaxPool2d((1, 1))
        self.out_dim = 512

    def _fwd(self, x, training, math, taps=None):
        if math == ops.MATH_FP32:
            xc = Act(ops.nchw_to_nhwc(x, c_pad=pad_channels(x.shape[1])).unsqueeze(1))   # (B, 1, T, F, 4): a 2-D layer is t == 1
            h, s_stem = ConvBNReLU.forward(xc, self.conv1[0], self.conv1[1], training, ops.MATH_FP32)  # CUDA-core kernel
        else:   # Cin = 1: the Toeplitz-view tcgen05 stem kernel
            op = StemOp(self.conv1[0], x.shape, math)
            h, s_stem = ConvBNReLU.forward(op.pack(x), self.conv1[0], self.conv1[1], training, math, op=op)
        saved_blocks = []
        for blk, tag in zip((self.block1, self.block2, self.block3, self.block4), ('conv2x', 'conv3x', 'conv4x', 'conv5x')):
            h, sb = blk._fwd(h, training, math)
            saved_blocks.append((blk, sb))
            if taps is not None:
                taps[tag] = h.f32
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
        ConvBNReLU.backward(d, s_stem, grads, need_dx=False, sums=sums)
