import sys, torch
sys.path.insert(0, "/root/repo")
from avid_cma_b200 import ops
n, ci, co, (t, h, w), k, s, p = 64, 3, 64, (8, 224, 224), (3, 7, 7), (1, 2, 2), (1, 3, 3)
x = torch.randn(n, ci, t, h, w, device="cuda")
wt = torch.randn(co, ci, *k, device="cuda") * 0.05
shape = ops.conv_shape(n, t, h, w, ci, co, k, s, p)
x_hi, x_lo = ops.stem_pack(x, 2 * shape.wo + 8, p[2])
w_hi, w_lo = ops.stem_filter_pack(wt)
out = torch.empty(n, shape.to, shape.ho, shape.wo, co, device="cuda")
for _ in range(2): ops.stem_forward_tc(shape, x_hi, x_lo, w_hi, w_lo, out=out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ops.stem_forward_tc(shape, x_hi, x_lo, w_hi, w_lo, out=out)
e1.record(); torch.cuda.synchronize()
print("stem fwd ms", e0.elapsed_time(e1) / 5)
