#!/usr/bin/env python
"""Timeline of one nce_gather_kernel launch from %globaltimer stamps (AVID_NCE_DEBUG=1): per-CTA start, first rows requested,
main loop done, ticket taken, query finalised, last-query tail.  Diagnostic only."""
import os
import sys
os.environ["AVID_NCE_DEBUG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from avid_cma_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
N, B = 2000000, 64
g = torch.Generator(device=dev).manual_seed(1)
bank_v = ops.rows_l2_normalize_(torch.randn(N, 128, device=dev, generator=g))
bank_a = ops.rows_l2_normalize_(torch.randn(N, 128, device=dev, generator=g))
ev, ea = torch.randn(B, 128, device=dev, generator=g), torch.randn(B, 128, device=dev, generator=g)
Z = torch.full((), 2.0, device=dev)
for K in [int(k) for k in sys.argv[1:]] or [256, 1024, 4096]:
    keys = [(0, 1, 0, K, 0.5), (1, 0, 0, K, 0.5)]
    ws = ops.nce_workspace(B, K, 0, 2, dev)
    out = torch.empty(3 + 2 * B * 128, device=dev)
    for it in range(4):
        y = torch.randint(0, N, (B,), device=dev, generator=g)
        args = ops.make_nce_args(ev, ea, y, bank_v, bank_a, keys, K, Z, seed=1, offset=it * B * K, loss_keys=out[1:3], loss_total=out[0:1],
                                 grad_v=out[3:3 + B * 128].view(B, 128), grad_a=out[3 + B * 128:].view(B, 128))
        ws.zero_()
        torch.cuda._sleep(200000)
        ops.nce_forward_backward(args, ws)
        torch.cuda.synchronize()
    off = ((4 * (B + 1) + 255) // 256) * 256
    st = ws.view(torch.uint8)[off:off + 1024 * 64].view(torch.int64).view(1024, 8).cpu()
    used = st[:, 0] > 0
    st = st[used]
    t0 = int(st[:, 0].min())
    names = ["start", "rows requested", "loop done", "ticket", "query finalised", "2nd ticket", "end of last query"]
    print(f"K={K}: {int(used.sum())} CTAs")
    for i, n in enumerate(names):
        col = st[:, i]
        col = col[col > 0] - t0
        if len(col):
            print(f"  {n:20s} n={len(col):4d} min {col.min() / 1e3:7.2f} us  median {col.median() / 1e3:7.2f}  max {col.max() / 1e3:7.2f}")
