#!/usr/bin/env python
"""Stand-alone timing of the BatchNorm elementwise passes on the conv2x tensor of config 2 (64 x 8 x 56 x 56 x 64 fp32 = 411 MB):
algorithmic bytes / time against the measured HBM copy rate."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from avid_cma_b200 import ops

DEV = "cuda:0"


def time_it(fn, iters=20):
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
    for shape in [(64, 8, 56, 56, 64), (64, 4, 28, 28, 128)]:
        n = 1
        for d in shape:
            n *= d
        c = shape[-1]
        x = torch.randn(shape, device=DEV)
        dy = torch.randn(shape, device=DEV)
        gamma, beta = torch.rand(c, device=DEV) + 0.5, torch.randn(c, device=DEV) * 0.1
        st = ops.BNState(c, DEV)
        st.mean.normal_(0, 0.1); st.invstd.fill_(1.0); st.scale.copy_(gamma); st.shift.copy_(beta)
        sums = torch.randn(2, c, device=DEV, dtype=torch.float64)
        print(shape, "%.0f MB per fp32 tensor" % (n * 4 / 1e6))
        for label, fn, nbytes in [
            ("forward  -> planes", lambda: ops.bn_relu_forward_act(x, st.scale, st.shift, False, True, True), n * 8),
            ("forward  -> planes + fp32", lambda: ops.bn_relu_forward_act(x, st.scale, st.shift, True, True, True), n * 12),
            ("backward -> planes", lambda: ops.bn_relu_backward_act(x, dy, st, gamma, beta, False, True, True, sums=sums), n * 12),
            ("backward -> planes + fp32", lambda: ops.bn_relu_backward_act(x, dy, st, gamma, beta, True, True, True, sums=sums), n * 16)]:
            us = time_it(fn)
            print("   %-28s %8.1f us  %7.0f GB/s  (%.2f of 6546)" % (label, us, nbytes / us / 1e3, nbytes / us / 1e3 / 6546), flush=True)


if __name__ == "__main__":
    main()
