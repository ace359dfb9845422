import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import torch.nn.functional as F
from avid_cma_b200 import ops
DEV='cuda:0'
n, ci, (t, h, w), k, s, p = (2, 3, (4, 20, 36), (3, 7, 7), (1, 2, 2), (1, 3, 3))
co=64
g = torch.Generator().manual_seed(1)
x = torch.randn(n, ci, t, h, w, generator=g)
wt = torch.randn(co, ci, *k, generator=g) / (ci * k[0] * k[1] * k[2]) ** 0.5
shape = ops.conv_shape(n, t, h, w, ci, co, k, s, p)
hp = ops.stem_packed_rows(shape)
print('hp', hp, 'wo', shape.wo, 'ho', shape.ho)
x_hi, x_lo = ops.stem_pack(x.to(DEV), 2 * shape.wo + 8, p[2], True, hp=hp, pad_top=p[1] % 2)
torch.cuda.synchronize(); print('pack ok')
w_hi, w_lo = ops.stem_filter_pack(wt.to(DEV), True)
torch.cuda.synchronize(); print('filter pack ok')
out = ops.stem_forward_tc(shape, x_hi, x_lo, w_hi, w_lo)
torch.cuda.synchronize(); print('forward ok')
ref = F.conv3d(x.double(), wt.double(), stride=s, padding=p)
got = ops.nhwc_to_nchw(out).cpu().double()
print('fwd rel err', float((got-ref).norm()/ref.norm()))
dout = torch.randn(ref.shape, generator=g)
d_hi, d_lo = ops.split_bf16(ops.nchw_to_nhwc(dout.to(DEV)), True)
dw_tap = ops.stem_wgrad_tc(shape, x_hi, x_lo, d_hi, d_lo)
torch.cuda.synchronize(); print('wgrad ok')
xd, wd = x.double(), wt.double().requires_grad_(True)
F.conv3d(xd, wd, stride=s, padding=p).backward(dout.double())
dw = ops.filter_from_tapmajor(dw_tap, wt.to(DEV)).cpu().double()
print('wgrad rel err', float((dw-wd.grad).norm()/wd.grad.norm()))
