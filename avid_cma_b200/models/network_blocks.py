"""Building blocks of the encoders on the CUDA kernels (reference: models/network_blocks.py).

The nn.Conv*/nn.BatchNorm*/nn.Linear children are PARAMETER CONTAINERS only: they give every module
the reference's state_dict keys, shapes and default initialisation (so published checkpoints load and
a seeded construction reproduces the reference's weights), but their own forward is never called.
All arithmetic is done by libavid_b200.so on channels-last activations [n, t, h, w, c]; a 2-D layer is
the t == 1 case.  Backward is hand-written (no autograd graph inside a tower): every block returns
the tensors its backward needs in a `saved` record.
"""
import torch
import torch.nn as nn

from .. import ops


def _triple(v, fill=1):
    """(t, h, w) view of a Conv2d / Conv3d hyper-parameter; `fill` is the value of the missing t entry
    (1 for kernel_size / stride, 0 for padding)."""
    if isinstance(v, int):
        return (fill, v, v)
    v = tuple(v)
    return v if len(v) == 3 else (fill,) + v


def pad_channels(c):
    """Channel count the conv kernels accept: 4, 8 or a multiple of 16."""
    if c <= 4:
        return 4
    if c <= 8:
        return 8
    return (c + 15) // 16 * 16


class ConvBNReLU:
    """conv -> train/eval BatchNorm -> ReLU, with an optional residual addend fused into the conv epilogue."""

    @staticmethod
    def forward(x, conv, bn, training, math, addend=None):
        n, t, h, w, ci = x.shape
        k, s, p = _triple(conv.kernel_size), _triple(conv.stride), _triple(conv.padding, 0)
        shape = ops.conv_shape(n, t, h, w, ci, conv.out_channels, k, s, p)
        w_tap, w_tap_t = ops.filter_to_tapmajor(conv.weight.detach(), ci_pad=ci)
        z = ops.conv_forward(shape, x, w_tap, addend=addend, math=math, ci_real=conv.in_channels)
        if training:
            st = ops.bn_train_stats(z, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, bn.eps, bn.momentum)
            bn.num_batches_tracked += 1
        else:
            st = ops.BNState(conv.out_channels, x.device)
            st.invstd.copy_(torch.rsqrt(bn.running_var + bn.eps))
            st.mean.copy_(bn.running_mean)
            st.scale.copy_(bn.weight.detach() * st.invstd)
            st.shift.copy_(bn.bias.detach() - bn.running_mean * st.scale)
        y = ops.bn_relu_forward(z, st.scale, st.shift)
        return y, (shape, x, w_tap_t, z, st, conv, bn)

    @staticmethod
    def backward(dy, saved, grads, math, need_dx=True, dx_addend=None):
        """Returns (dx or None, dz) where dz is the gradient at the conv output (after the residual sum)."""
        shape, x, w_tap_t, z, st, conv, bn = saved
        dz, dgamma, dbeta = ops.bn_relu_backward(z, dy, st, bn.weight.detach(), bn.bias.detach())
        grads[bn.weight], grads[bn.bias] = dgamma, dbeta
        grads[conv.weight] = ops.filter_from_tapmajor(ops.conv_wgrad(shape, x, dz, math=math, ci_real=conv.in_channels), conv.weight)
        dx = ops.conv_dgrad(shape, dz, w_tap_t, addend=dx_addend, math=math) if need_dx else None
        return dx, dz


class Basic2DBlock(nn.Module):
    """(3x3 conv -> BN -> ReLU) x 2, no residual (network_blocks.py:13-27)."""

    def __init__(self, in_planes, out_planes, stride=(1, 1)):
        super().__init__()
        self.conv1 = nn.Conv2d(in_planes, out_planes, kernel_size=(3, 3), padding=(1, 1), stride=stride, bias=False)
        self.bn1 = nn.BatchNorm2d(out_planes)
        self.conv2 = nn.Conv2d(out_planes, out_planes, kernel_size=(3, 3), padding=(1, 1), bias=False)
        self.bn2 = nn.BatchNorm2d(out_planes)
        self.relu = nn.ReLU(inplace=True)

    def _fwd(self, x, training, math):
        y1, s1 = ConvBNReLU.forward(x, self.conv1, self.bn1, training, math)
        y2, s2 = ConvBNReLU.forward(y1, self.conv2, self.bn2, training, math)
        return y2, (s1, s2)

    def _bwd(self, dy, saved, grads, math, need_dx=True):
        s1, s2 = saved
        d1, _ = ConvBNReLU.backward(dy, s2, grads, math)
        dx, _ = ConvBNReLU.backward(d1, s1, grads, math, need_dx=need_dx)
        return dx


class BasicR2P1DBlock(nn.Module):
    """spt(1x3x3) -> tmp(3x1x1) -> spt -> tmp (+ identity or 1x1x1 strided conv WITHOUT BN) -> out_bn -> ReLU
    (network_blocks.py:30-60)."""

    def __init__(self, in_planes, out_planes, stride=(1, 1, 1)):
        super().__init__()
        spt_stride = (1, stride[1], stride[2])
        tmp_stride = (stride[0], 1, 1)
        self.spt_conv1 = nn.Conv3d(in_planes, out_planes, kernel_size=(1, 3, 3), stride=spt_stride, padding=(0, 1, 1), bias=False)
        self.spt_bn1 = nn.BatchNorm3d(out_planes)
        self.tmp_conv1 = nn.Conv3d(out_planes, out_planes, kernel_size=(3, 1, 1), stride=tmp_stride, padding=(1, 0, 0), bias=False)
        self.tmp_bn1 = nn.BatchNorm3d(out_planes)
        self.spt_conv2 = nn.Conv3d(out_planes, out_planes, kernel_size=(1, 3, 3), stride=(1, 1, 1), padding=(0, 1, 1), bias=False)
        self.spt_bn2 = nn.BatchNorm3d(out_planes)
        self.tmp_conv2 = nn.Conv3d(out_planes, out_planes, kernel_size=(3, 1, 1), stride=(1, 1, 1), padding=(1, 0, 0), bias=False)
        self.out_bn = nn.BatchNorm3d(out_planes)
        self.relu = nn.ReLU(inplace=True)
        if in_planes != out_planes or any([s != 1 for s in stride]):
            self.res = True
            self.res_conv = nn.Conv3d(in_planes, out_planes, kernel_size=(1, 1, 1), stride=stride, padding=(0, 0, 0), bias=False)
        else:
            self.res = False

    def _fwd(self, x, training, math):
        y1, s1 = ConvBNReLU.forward(x, self.spt_conv1, self.spt_bn1, training, math)
        y2, s2 = ConvBNReLU.forward(y1, self.tmp_conv1, self.tmp_bn1, training, math)
        y3, s3 = ConvBNReLU.forward(y2, self.spt_conv2, self.spt_bn2, training, math)
        sres = None
        if self.res:
            n, t, h, w, ci = x.shape
            rc = self.res_conv
            rshape = ops.conv_shape(n, t, h, w, ci, rc.out_channels, _triple(rc.kernel_size), _triple(rc.stride), _triple(rc.padding, 0))
            rw, rw_t = ops.filter_to_tapmajor(rc.weight.detach(), ci_pad=ci)
            r = ops.conv_forward(rshape, x, rw, math=math)
            sres = (rshape, rw_t)
        else:
            r = x
        # x_main + x_res is formed in the epilogue of tmp_conv2, then out_bn + ReLU (network_blocks.py:58-59)
        y4, s4 = ConvBNReLU.forward(y3, self.tmp_conv2, self.out_bn, training, math, addend=r)
        return y4, (s1, s2, s3, s4, sres, x)

    def _bwd(self, dy, saved, grads, math, need_dx=True):
        s1, s2, s3, s4, sres, x = saved
        d3, d_sum = ConvBNReLU.backward(dy, s4, grads, math)
        d2, _ = ConvBNReLU.backward(d3, s3, grads, math)
        d1, _ = ConvBNReLU.backward(d2, s2, grads, math)
        if self.res:
            rshape, rw_t = sres
            grads[self.res_conv.weight] = ops.filter_from_tapmajor(ops.conv_wgrad(rshape, x, d_sum, math=math), self.res_conv.weight)
            d_res = ops.conv_dgrad(rshape, d_sum, rw_t, math=math) if need_dx else None
        else:
            d_res = d_sum
        dx, _ = ConvBNReLU.backward(d1, s1, grads, math, need_dx=need_dx, dx_addend=d_res)
        return dx
