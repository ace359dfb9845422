#!/usr/bin/env python
"""Phase timing of the tensor-core CMA mining (to_half / scan_tc / rescore / certify / finish) on random unit rows."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avid_cma_b200 import ops, _lib
from avid_cma_b200.ops import _p, _stream, check

N = int(sys.argv[1]) if len(sys.argv) > 1 else 240000
dev = "cuda:0"
torch.manual_seed(0)
bv = torch.nn.functional.normalize(torch.randn(N, 128, device=dev), dim=1)
ba = torch.nn.functional.normalize(torch.randn(N, 128, device=dev), dim=1)
L = _lib.lib()
ws = torch.empty(int(L.avid_cma_topk_workspace_bytes(N)), dtype=torch.uint8, device=dev)
wp, wn = _p(ws, torch.uint8), ws.numel()


def timed(name, fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    torch.cuda.synchronize()
    print("%-10s %9.2f ms" % (name, e0.elapsed_time(e1)))
    return r


for rep in range(2):
    timed("begin", lambda: check(L.avid_cma_topk_begin(N, wp, wn, _stream())))
    qv, qa = timed("to_half", lambda: (ops._cma_to_half(bv), ops._cma_to_half(ba)))
    for mode in ([0] if rep == 0 else [0, 2]):
        timed("begin", lambda: check(L.avid_cma_topk_begin(N, wp, wn, _stream())))
        timed("scan_tc m%d" % mode, lambda: check(L.avid_cma_topk_scan_tc(_p(qv, torch.float16), _p(qa, torch.float16), N, _p(qv, torch.float16), _p(qa, torch.float16), 0, N, mode, wp, wn, _stream())))
    timed("rescore", lambda: check(L.avid_cma_topk_rescore(_p(bv), _p(ba), N, _p(bv), _p(ba), 0, N, 0, wp, wn, _stream())))
    fc = torch.zeros(1, dtype=torch.int32, device=dev)
    fl = torch.empty(N, dtype=torch.int32, device=dev)
    timed("certify", lambda: check(L.avid_cma_topk_certify(N, 32, 1e-3, wp, wn, _p(fc, torch.int32), _p(fl, torch.int32), _stream())))
    out = torch.empty(N, 32, dtype=torch.int32, device=dev)
    timed("finish", lambda: check(L.avid_cma_topk_finish(N, 32, _p(out, torch.int32), wp, wn, _stream())))
    print("uncertified", int(fc.item()))
