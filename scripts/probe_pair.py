#!/usr/bin/env python
"""Where does a conv_pair launch (64 -> 64 channel 1x3x3 layer on CTA pairs) spend its time?  AVID_PAIR_DEBUG bits switch parts off:
1 no epilogue stores, 8 no activation loads, 256 report the launch geometry."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avid_cma_b200 import ops

DEV = "cuda:0"


def time_it(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    n, ci, co, (t, h, w), k, p = 64, 64, 64, (8, 56, 56), (1, 3, 3), (0, 1, 1)
    shape = ops.conv_shape(n, t, h, w, ci, co, k, (1, 1, 1), p)
    x = torch.randn(n, t, h, w, ci, device=DEV)
    wt = torch.randn(9, co, ci, device=DEV) / (ci * 9) ** 0.5
    x_hi, x_lo = ops.split_bf16(x, True)
    w_hi, w_lo = ops.split_bf16(wt, True)
    out = torch.empty(n, t, h, w, co, device=DEV)
    stats = torch.zeros(2, co, dtype=torch.float64, device=DEV)
    flops = 2.0 * n * t * h * w * ci * co * 9
    for label, dbg in [("full", 0), ("no stores", 1), ("no A loads", 8), ("no A loads, no stores", 9), ("full + report", 256)]:
        os.environ["AVID_PAIR_DEBUG"] = str(dbg)
        us = time_it(lambda: ops.conv_forward_tc(shape, x_hi, x_lo, w_hi, w_lo, out=out, bn_stats=stats), iters=1 if dbg & (256 | 512) else 10)
        print("   %-36s %8.1f us  %7.1f TFLOP/s" % (label, us, flops / us / 1e6), flush=True)
    os.environ["AVID_PAIR_DEBUG"] = "0"
    # input gradient (same kernel, mirrored taps) with the residual addend and the fused BatchNorm-backward sums
    wt_t = torch.randn(9, ci, co, device=DEV) / (ci * 9) ** 0.5
    wt_hi, wt_lo = ops.split_bf16(wt_t, True)
    add = torch.randn(n, t, h, w, ci, device=DEV)
    z = torch.randn(n, t, h, w, ci, device=DEV)
    st = ops.BNState(ci, DEV)
    st.mean.normal_(); st.invstd.fill_(1.0)
    gamma, beta = torch.ones(ci, device=DEV), torch.zeros(ci, device=DEV)
    sums = torch.zeros(2, ci, dtype=torch.float64, device=DEV)
    for label, kw in [("dgrad", {}), ("dgrad + addend", dict(addend=add)), ("dgrad + fused BN sums", dict(bn_fuse=(z, st, gamma, beta, sums))),
                      ("dgrad + addend + fused BN sums", dict(addend=add, bn_fuse=(z, st, gamma, beta, sums)))]:
        us = time_it(lambda: ops.conv_dgrad_tc(shape, x_hi, x_lo, wt_hi, wt_lo, out=out, **kw))
        print("   %-36s %8.1f us  %7.1f TFLOP/s" % (label, us, flops / us / 1e6), flush=True)
    us = time_it(lambda: ops.conv_forward_tc(shape, x_hi, None, w_hi, None, out=out, bn_stats=stats))
    print("   %-36s %8.1f us  %7.1f TFLOP/s" % ("bf16 single pass", us, flops / us / 1e6), flush=True)


if __name__ == "__main__":
    main()
