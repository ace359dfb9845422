import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import criterion as oc, synth
from avid_cma_b200 import ops
DEV = "cuda:0"

def run(N, B, K, xw):
    bv, ba = synth.bank(N, seed=21, tag="bank_v"), synth.bank(N, seed=21, tag="bank_a")
    ev, ea = synth.embeddings(B, seed=21)
    y = synth.instance_ids(B, N, seed=21)
    idx = synth.negatives(y, K, N, seed=21)
    keys = oc.avid_keys(K, *xw)
    kt = [({"v": 0, "a": 1}[k.ctx], {"v": 0, "a": 1}[k.bank], 0, k.num_neg, k.weight) for k in keys]
    d = lambda t: t.to(DEV)
    Z = torch.tensor(1.7, device=DEV)
    nk = len(keys)
    out = [torch.empty(nk, device=DEV), torch.empty(1, device=DEV), torch.empty(B, 128, device=DEV), torch.empty(B, 128, device=DEV)]
    scores = torch.full((nk, B, 1 + K), float("nan"), device=DEV)
    lp = torch.empty(nk, B, device=DEV)
    a = ops.make_nce_args(d(ev), d(ea), d(y), d(bv), d(ba), kt, K, Z, neg_idx=d(idx), loss_keys=out[0], loss_total=out[1],
                          grad_v=out[2], grad_a=out[3], scores=scores, loss_part=lp)
    ops.nce_forward_backward(a, ops.nce_workspace(B, K, 0, nk, DEV))
    torch.cuda.synchronize()
    r = oc.criterion_forward_backward(ev, ea, y, bv, ba, idx, keys, 1.7, dtype=torch.float64)
    sc = oc.scores(ev.double(), ea.double(), y, bv, ba, idx, keys)
    print(f"--- N={N} B={B} K={K} keys={[k.name for k in keys]}")
    print("total", float(out[1]), float(r["total"]))
    for i, k in enumerate(keys):
        sp, sn = sc[k.name]
        mine = scores[i].cpu().double()
        nan = int(torch.isnan(mine).sum())
        dpos = float((mine[:, :1] - sp).abs().nan_to_num(99).max()); dneg = float((mine[:, 1:] - sn).abs().nan_to_num(99).max())
        # per-instance loss terms from oracle scores
        c = K * 1.7
        per_b = torch.log1p(c / torch.exp(sp)).mean(1) + torch.log1p(torch.exp(sn) / c).sum(1)
        dl = (lp[i].cpu().double() - per_b).abs()
        print(k.name, "loss", float(out[0][i]), float(r["losses"][k.name]), "nan scores", nan, "dpos", dpos, "dneg", dneg,
              "max |loss_part diff|", float(dl.max()), "argmax b", int(dl.argmax()), "mean diff", float((lp[i].cpu().double() - per_b).mean()))
    print("grad_v rel", float((out[2].cpu().double() - r["grad_v"]).norm() / r["grad_v"].norm()))

run(240000, 64, 1024, (1.0, 1.0))
run(240000, 64, 1024, (1.0, 0.0))
run(240000, 4, 1024, (1.0, 1.0))
run(300, 64, 64, (1.0, 1.0))
run(50000, 33, 100, (1.0, 1.0))
