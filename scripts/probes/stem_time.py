"""Time (CUDA events) the two stem kernels at the BASELINE config-2 video shape; run under ncu for profiles."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from avid_cma_b200 import ops

DEV = "cuda:0"
n, ci, (t, h, w), k, s, p = 64, 3, (8, 224, 224), (3, 7, 7), (1, 2, 2), (1, 3, 3)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
x = torch.randn(n, ci, t, h, w, device=DEV)
wt = torch.randn(64, ci, *k, device=DEV) / 21.0
shape = ops.conv_shape(n, t, h, w, ci, 64, k, s, p)
x_hi, x_lo = ops.stem_pack(x, 2 * shape.wo + 8, p[2], True)
w_hi, w_lo = ops.stem_filter_pack(wt, True)
out = torch.empty(n, shape.to, shape.ho, shape.wo, 64, device=DEV)
stats = torch.zeros(2, 64, dtype=torch.float64, device=DEV)
dz = torch.randn(n, shape.to, shape.ho, shape.wo, 64, device=DEV)
d_hi, d_lo = ops.split_bf16(dz, True)
del dz
for name, fn in (("stem_forward", lambda: ops.stem_forward_tc(shape, x_hi, x_lo, w_hi, w_lo, out=out, bn_stats=stats)),
                 ("stem_wgrad", lambda: ops.stem_wgrad_tc(shape, x_hi, x_lo, d_hi, d_lo))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print("%s %.3f ms  %.1f TFLOP/s" % (name, ms, 2.0 * n * shape.to * shape.ho * shape.wo * 441 * 64 / ms / 1e9), flush=True)
