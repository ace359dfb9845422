#!/usr/bin/env python
"""Where does a conv_tc launch spend its time?  Times single layers of the config-2 video tower with parts of the kernel
switched off through AVID_TC_DEBUG (1 no global stores, 2 no statistics, 4 no epilogue work, 8 no MMAs)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avid_cma_b200 import ops

DEV = "cuda:0"
LAYERS = [("tmp 64->64 8x56x56", 64, 64, (8, 56, 56), (3, 1, 1), (1, 0, 0)),
          ("spt 64->64 8x56x56", 64, 64, (8, 56, 56), (1, 3, 3), (0, 1, 1)),
          ("tmp 128->128 4x28x28", 128, 128, (4, 28, 28), (3, 1, 1), (1, 0, 0))]


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
    n = 64
    for name, ci, co, (t, h, w), k, p in LAYERS:
        shape = ops.conv_shape(n, t, h, w, ci, co, k, (1, 1, 1), p)
        x = torch.randn(n, t, h, w, ci, device=DEV)
        wt = torch.randn(k[0] * k[1] * k[2], co, ci, device=DEV) / (ci * k[0] * k[1] * k[2]) ** 0.5
        x_hi, x_lo = ops.split_bf16(x, True)
        w_hi, w_lo = ops.split_bf16(wt, True)
        out = torch.empty(n, t, h, w, co, device=DEV)
        add = torch.randn(n, t, h, w, co, device=DEV)
        stats = torch.zeros(2, co, dtype=torch.float64, device=DEV)
        flops = 2.0 * n * t * h * w * ci * co * k[0] * k[1] * k[2]
        print(name)
        for label, dbg, kw in [("full (stats)", 0, dict(bn_stats=stats)), ("full + addend", 0, dict(bn_stats=stats, addend=add)), ("no stats arg", 0, {}),
                               ("no global stores", 1, dict(bn_stats=stats)), ("no stores, no stats", 3, dict(bn_stats=stats)),
                               ("no epilogue work", 4, dict(bn_stats=stats)), ("no MMAs", 8, dict(bn_stats=stats)),
                               ("no MMAs, no epilogue", 12, dict(bn_stats=stats)), ("bf16 single pass", 0, dict(bn_stats=stats, single=True)),
                               ("no atomics block", 64, dict(bn_stats=stats)), ("noMMA", 8, dict(bn_stats=stats)), ("noMMA no atomics", 72, dict(bn_stats=stats)),
                               ("noMMA noatom nostore", 73, dict(bn_stats=stats)), ("noMMA noatom nostore nostat", 75, dict(bn_stats=stats))]:
            os.environ["AVID_TC_DEBUG"] = str(dbg)
            single = kw.pop("single", False)
            us = time_it(lambda: ops.conv_forward_tc(shape, x_hi, None if single else x_lo, w_hi, None if single else w_lo, out=out, **kw))
            print("   %-22s %8.1f us  %7.1f TFLOP/s" % (label, us, flops / us / 1e6))
        os.environ["AVID_TC_DEBUG"] = "0"


if __name__ == "__main__":
    main()
