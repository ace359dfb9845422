"""Building blocks of the encoders on the CUDA kernels (reference: models/network_blocks.py).

The nn.Conv*/nn.BatchNorm*/nn.Linear children are PARAMETER CONTAINERS only: they give every module
the reference's state_dict keys, shapes and default initialisation (so published checkpoints load and
a seeded construction reproduces the reference's weights), but their own forward is never called.
All arithmetic is done by libavid_b200.so on channels-last activations [n, t, h, w, c]; a 2-D layer is
the t == 1 case.  Backward is hand-written (no autograd graph inside a tower): every block returns
the tensors its backward needs in a `saved` record.

Math modes (ops.MATH_*): FP32 runs the CUDA-core implicit-GEMM kernels; BF16X3 / BF16 run every layer
whose channel counts are multiples of 64 (all but the two stems) on the tcgen05 kernels, with
activations and gradients handed from layer to layer as bf16 (hi, lo) planes written by the
BatchNorm+ReLU kernels.  Strided input gradients run as one tcgen05 launch per stride-parity class.
"""
import os

import torch
import torch.nn as nn

from .. import ops
from ..ops import Act


FUSE_MIN_K = int(os.environ.get('AVID_FUSE_BN_BWD_MIN_K', '192'))    # contraction length (co * taps) from which dgrad also reduces BN backward


_BATCH_COUNTERS = []


def flush_batch_counters():
    """num_batches_tracked += 1 of every BatchNorm the tower just ran in training mode (nn.BatchNorm's bookkeeping), as ONE
    multi-tensor launch instead of a tiny kernel per layer."""
    if _BATCH_COUNTERS:
        torch._foreach_add_(list(_BATCH_COUNTERS), 1)
        del _BATCH_COUNTERS[:]


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


class ConvOp:
    """One convolution of a step: geometry + the filter in every layout / precision its kernels read."""

    def __init__(self, conv, x_shape, math):
        n, t, h, w, ci = x_shape
        self.conv = conv
        self.k, self.s, self.p = _triple(conv.kernel_size), _triple(conv.stride), _triple(conv.padding, 0)
        self.shape = ops.conv_shape(n, t, h, w, ci, conv.out_channels, self.k, self.s, self.p)
        self.tc = math != ops.MATH_FP32 and ci % 64 == 0 and conv.out_channels % 64 == 0
        self.x3 = math == ops.MATH_BF16X3
        self.unit_stride = self.s == (1, 1, 1)
        self.ci_real = conv.in_channels
        if self.tc:
            # forward operand [taps, co, ci] (K-major in ci) and dgrad operand [taps, ci, co] (K-major in co), one launch
            (self.wf_hi, self.wf_lo), (self.wd_hi, self.wd_lo) = ops.filter_to_planes(conv.weight.detach(), self.x3)
            self.w_tap = self.w_tap_t = None
        else:
            self.w_tap, self.w_tap_t = ops.filter_to_tapmajor(conv.weight.detach(), ci_pad=ci)   # fp32 [taps, ci, co] / [taps, co, ci]

    def forward(self, x, addend=None, bn_stats=None):
        """bn_stats: (2, co) fp64 zeros the tensor-core epilogue accumulates the BatchNorm statistics into (tc only)."""
        if self.tc:
            x.ensure_planes(self.x3)
            return ops.conv_forward_tc(self.shape, x.hi, x.lo, self.wf_hi, self.wf_lo, addend=addend, ci_real=self.ci_real, bn_stats=bn_stats)
        return ops.conv_forward(self.shape, x.f32, self.w_tap, addend=addend, ci_real=self.ci_real)

    def needs_f32_dz(self):
        return not self.tc

    def wgrad(self, x, dz):
        if self.tc:
            dw = ops.conv_wgrad_tc(self.shape, x.hi, x.lo, dz.hi, dz.lo, ci_real=self.ci_real)
        else:
            dw = ops.conv_wgrad(self.shape, x.f32, dz.f32, ci_real=self.ci_real)
        return ops.filter_from_tapmajor(dw, self.conv.weight)

    def can_fuse_bn_backward(self):
        """The tensor-core input gradient can also reduce the previous layer's BatchNorm backward (every input pixel must be
        written by a computed tile: filter >= stride)."""
        # Measured on B200 (scripts/gpu_fuse_sweep.sh): with the 8-warp epilogue of conv_tc_kernel the fusion pays for every layer
        # with K >= 192 (input-gradient launches 4.8 -> 5.7 ms per step, but 1.9 ms of separate reduction passes disappear:
        # 2123 -> 2176 clips/s); only the 1x1x1 residual convolutions (K = 64..256 with one tap) stay unfused.  With the earlier
        # 4-warp epilogue the same fusion made the 64-channel launches epilogue-bound (5.3 -> 9.7 ms) and was limited to K >= 2304.
        k_total = self.conv.out_channels * self.k[0] * self.k[1] * self.k[2]
        return self.tc and all(k >= s for k, s in zip(self.k, self.s)) and k_total >= FUSE_MIN_K

    def dgrad(self, dz, addend=None, bn_fuse=None, addend_stride=None):
        if self.tc:
            return ops.conv_dgrad_tc(self.shape, dz.hi, dz.lo, self.wd_hi, self.wd_lo, addend=addend, bn_fuse=bn_fuse, addend_stride=addend_stride)
        assert bn_fuse is None and addend_stride is None
        return ops.conv_dgrad(self.shape, dz.f32, self.w_tap_t, addend=addend)

    def subsampled_dgrad(self, dz):
        """Strided 1x1x1 convolution on the tensor cores: its input gradient at the pixels it reads, [n, to, ho, wo, ci] (a
        stride-1 1x1x1 input gradient over the OUTPUT grid; every other input pixel receives zero) and the stride, for
        ops.conv_dgrad_tc(addend_stride=...).  None when the layer is not such a convolution."""
        if not (self.tc and self.k == (1, 1, 1) and self.p == (0, 0, 0) and self.s != (1, 1, 1)):
            return None
        sh = self.shape
        dense = ops.conv_shape(sh.n, sh.to, sh.ho, sh.wo, sh.ci, sh.co, (1, 1, 1), (1, 1, 1), (0, 0, 0))
        return ops.conv_dgrad_tc(dense, dz.hi, dz.lo, self.wd_hi, self.wd_lo), self.s


class StemOp:
    """The first convolution of a tower (Cin = 3 / 1, 7x7 stride 2) on the tcgen05 stem kernels (csrc/stem_tc.cu): same
    interface as ConvOp, but its input Act holds the packed planes [n, t, h, wp, 4] written by `pack`."""
    tc = True
    unit_stride = False

    def __init__(self, conv, x_shape, math):
        n, c = x_shape[0], x_shape[1]
        t, h, w = (1,) * (5 - len(x_shape)) + tuple(x_shape[2:])
        self.conv = conv
        self.k, self.s, self.p = _triple(conv.kernel_size), _triple(conv.stride), _triple(conv.padding, 0)
        self.shape = ops.conv_shape(n, t, h, w, c, conv.out_channels, self.k, self.s, self.p)
        self.x3 = math == ops.MATH_BF16X3
        self.wp = 2 * self.shape.wo + 8
        self.w_hi, self.w_lo = ops.stem_filter_pack(conv.weight.detach(), self.x3)

    def pack(self, x):
        hi, lo = ops.stem_pack(x, self.wp, self.p[2], self.x3)
        return Act(None, hi, lo)

    def forward(self, x, addend=None, bn_stats=None):
        assert addend is None
        return ops.stem_forward_tc(self.shape, x.hi, x.lo, self.w_hi, self.w_lo, bn_stats=bn_stats)

    def needs_f32_dz(self):
        return False

    def wgrad(self, x, dz):
        return ops.filter_from_tapmajor(ops.stem_wgrad_tc(self.shape, x.hi, x.lo, dz.hi, dz.lo), self.conv.weight)


class ConvBNReLU:
    """conv -> train/eval BatchNorm -> ReLU, with an optional residual addend fused into the conv epilogue."""

    @staticmethod
    def forward(x, conv, bn, training, math, addend=None, out_f32=False, op=None, out_planes=None, pool=False):
        """x: Act.  Returns (y: Act, saved).  In the tensor-core modes y carries the bf16 planes the next
        convolution reads (plus fp32 when `out_f32`: block outputs feed residual adds and pools).  `pool`: y is
        additionally max-pooled (1,3,3)/(1,2,2) in the same kernel (the video stem); the ReLU output is not stored."""
        if op is None:
            op = ConvOp(conv, x.shape, math)
        if training:
            # tensor-core convolutions accumulate the batch statistics in their epilogue; the fp32 kernels need a separate pass
            st = ops.BNState(conv.out_channels, x.device) if op.tc else None
            z = op.forward(x, addend, bn_stats=st.stats) if op.tc else op.forward(x, addend)
            st = ops.bn_train_stats(z, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, bn.eps, bn.momentum, state=st)
            _BATCH_COUNTERS.append(bn.num_batches_tracked)       # += 1 for all layers of the tower in one foreach launch
        else:
            z = op.forward(x, addend)
            st = ops.BNState(conv.out_channels, z.device)
            st.frozen = True
            st.invstd.copy_(torch.rsqrt(bn.running_var + bn.eps))
            st.mean.copy_(bn.running_mean)
            st.scale.copy_(bn.weight.detach() * st.invstd)
            st.shift.copy_(bn.bias.detach() - bn.running_mean * st.scale)
        planes = math != ops.MATH_FP32 if out_planes is None else out_planes
        if pool:
            y, amax = ops.bn_relu_maxpool_forward(z, st.scale, st.shift, want_f32=True, want_planes=planes, x3=math == ops.MATH_BF16X3)
            return y, (op, x, z, st, bn, amax, y.f32)
        y = ops.bn_relu_forward_act(z, st.scale, st.shift, want_f32=out_f32 or not planes, want_planes=planes, x3=math == ops.MATH_BF16X3)
        return y, (op, x, z, st, bn)

    @staticmethod
    def backward(dy, saved, grads, need_dx=True, dx_addend=None, dz_f32=False, sums=None, below=None, dx_addend_stride=None):
        """dy: fp32 gradient at the ReLU output.  Returns (dx fp32 or None, dz: Act at the conv output, i.e. after
        the residual sum, sums_below).

        `below`: the saved record of the conv-BN-ReLU layer whose output is this layer's input.  When given (and this layer's
        input gradient runs on the tensor cores) the dgrad epilogue also reduces THAT layer's BatchNorm backward; the (2, c)
        sums are returned as `sums_below` and must be passed as `sums` to that layer's backward, which then skips its own
        reduction pass over dy and z."""
        op, x, z, st, bn = saved[:5]
        want_f32 = dz_f32 or not op.tc or (need_dx and op.needs_f32_dz())
        if len(saved) > 5:      # pooled layer: dy is the gradient at the pooled output
            dz, dgamma, dbeta = ops.bn_relu_maxpool_backward_act(z, saved[6], saved[5], dy, st, bn.weight.detach(), bn.bias.detach(),
                                                                 want_f32=want_f32, want_planes=op.tc, x3=op.x3)
        else:
            dz, dgamma, dbeta = ops.bn_relu_backward_act(z, dy, st, bn.weight.detach(), bn.bias.detach(), want_f32=want_f32,
                                                         want_planes=op.tc, x3=op.x3, sums=sums)
        grads[bn.weight], grads[bn.bias] = dgamma, dbeta
        side = ops.wgrad_stream_enabled() and op.tc and need_dx
        if not side:
            grads[op.conv.weight] = op.wgrad(x, dz)
        dx, sums_below = None, None
        if need_dx:
            fuse = None
            if below is not None and len(below) == 5 and getattr(op, 'can_fuse_bn_backward', lambda: False)():
                _, _, z_b, st_b, bn_b = below
                sums_below = ops.bn_backward_sums(z_b.shape[-1], z_b.device)
                fuse = (z_b, st_b, bn_b.weight.detach(), bn_b.bias.detach(), sums_below)
            if dx_addend_stride is not None:
                dx = op.dgrad(dz, addend=dx_addend, bn_fuse=fuse, addend_stride=dx_addend_stride)
            else:
                dx = op.dgrad(dz, addend=dx_addend, bn_fuse=fuse) if fuse is not None else op.dgrad(dz, addend=dx_addend)
        if side:
            # The filter gradient starts when the input gradient has DRAINED (event after it, AVID_WGRAD_STREAM=2: before it): two persistent
            # tensor-core grids cannot share the SMs anyway, and this way the wgrad CTAs and the BatchNorm-backward pass of the next layer
            # down (main stream, no shared memory) start together instead of the wgrad winning the SMs ahead of the critical path
            ready = ops.side_event()
            grads[op.conv.weight] = ops.side_run(ready, lambda: op.wgrad(x, dz), keep=(x, dz))
        return dx, dz, sums_below


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
        y2, s2 = ConvBNReLU.forward(y1, self.conv2, self.bn2, training, math, out_f32=True)
        return y2, (s1, s2)

    def _bwd(self, dy, saved, grads, need_dx=True, sums=None, below=None):
        """`sums`: BatchNorm-backward sums of bn2 already reduced by the block above; `below`: record of the layer feeding this
        block.  Returns (dx, sums for `below`)."""
        s1, s2 = saved
        d1, _, sums1 = ConvBNReLU.backward(dy, s2, grads, sums=sums, below=s1)
        dx, _, sums_below = ConvBNReLU.backward(d1, s1, grads, need_dx=need_dx, sums=sums1, below=below)
        return dx, sums_below

    @staticmethod
    def top_record(saved):
        """The record of the block's last conv-BN-ReLU (the layer a block above may reduce the BatchNorm backward of)."""
        return saved[1]


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
        """x: Act carrying fp32 (the identity residual) and, in tensor-core modes, planes."""
        y1, s1 = ConvBNReLU.forward(x, self.spt_conv1, self.spt_bn1, training, math)
        y2, s2 = ConvBNReLU.forward(y1, self.tmp_conv1, self.tmp_bn1, training, math)
        y3, s3 = ConvBNReLU.forward(y2, self.spt_conv2, self.spt_bn2, training, math)
        rop = None
        if self.res:
            rop = ConvOp(self.res_conv, x.shape, math)
            r = rop.forward(x)
        else:
            r = x.f32
        # x_main + x_res is formed in the epilogue of tmp_conv2, then out_bn + ReLU (network_blocks.py:58-59)
        y4, s4 = ConvBNReLU.forward(y3, self.tmp_conv2, self.out_bn, training, math, addend=r, out_f32=True)
        return y4, (s1, s2, s3, s4, rop, x)

    def _bwd(self, dy, saved, grads, need_dx=True, sums=None, below=None):
        """`sums`: BatchNorm-backward sums of out_bn already reduced by the block above; `below`: record of the layer feeding
        this block.  Returns (dx, sums for `below`)."""
        s1, s2, s3, s4, rop, x = saved
        # d_sum (gradient at x_main + x_res) flows to tmp_conv2 and to the residual branch
        d3, d_sum, sums3 = ConvBNReLU.backward(dy, s4, grads, dz_f32=(not self.res) or rop.needs_f32_dz(), sums=sums, below=s3)
        d2, _, sums2 = ConvBNReLU.backward(d3, s3, grads, sums=sums3, below=s2)
        d1, _, sums1 = ConvBNReLU.backward(d2, s2, grads, sums=sums2, below=s1)
        res_stride = None
        if self.res:
            grads[self.res_conv.weight] = rop.wgrad(x, d_sum)
            d_res = None
            if need_dx:
                # a strided 1x1x1 residual branch: its input gradient stays on the output grid (no zero fill of the input grid, and
                # spt_conv1's input gradient reads 1/8 of the addend bytes)
                sub = rop.subsampled_dgrad(d_sum) if s1[0].tc and not isinstance(s1[0], StemOp) else None
                if sub is not None:
                    d_res, res_stride = sub
                else:
                    d_res = rop.dgrad(d_sum)
        else:
            d_res = d_sum.f32
        # the block's input gradient = spt_conv1's input gradient + the residual branch's (fused as the epilogue addend), so the
        # epilogue sees the complete gradient and can reduce the BatchNorm backward of the layer below
        dx, _, sums_below = ConvBNReLU.backward(d1, s1, grads, need_dx=need_dx, dx_addend=d_res, sums=sums1, below=below,
                                                dx_addend_stride=res_stride)
        return dx, sums_below

    @staticmethod
    def top_record(saved):
        return saved[3]
