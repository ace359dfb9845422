#!/usr/bin/env python
"""BASELINE.json config 4: AVID+CMA (InstX-N1024-PosW-N64-Top32) on one B200 -- (a) positive mining over the whole bank
(avid_cma_topk_*: consensus similarity, top-32), (b) the criterion step with the 4 score keys (inst-v2a/a2v K=1024,
pos-v2v/a2a P=32 K=64).  Mining FLOPs = 2 banks x 2 N^2 D.

    python scripts/bench_cma.py [--bank 240000] [--out profiles/r1_cma.json]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bank", type=int, default=240000)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    from avid_cma_b200.criterions import AVID_CMA
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    N, B = a.bank, a.batch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    crit = AVID_CMA(num_data=N, embedding_dim=128, num_negatives=1024, num_negatives_within=64, momentum=0.5,
                    xModalInstCoeff=1., wModalInstCoeff=0., xModalPosCoeff=0., wModalPosCoeff=1.,
                    sampling_args={"type": "consensus", "pos_k": 32}, device=0)      # mines in the constructor (avid_cma.py:322)
    e1.record()
    torch.cuda.synchronize()
    ctor_s = time.perf_counter() - t0
    e0.record()
    crit.nce_average.find_correspondences()
    e1.record()
    torch.cuda.synchronize()
    mine_ms = e0.elapsed_time(e1)
    flops = 2 * 2.0 * N * N * 128
    ps = crit.nce_average.positive_set
    assert ps.shape == (N, 32) and int(ps.min()) >= 0 and int(ps.max()) < N
    g = torch.Generator(device=dev).manual_seed(1)
    times = []
    for it in range(5 + a.iters):
        ev = torch.randn(B, 128, device=dev, generator=g, requires_grad=True)
        ea = torch.randn(B, 128, device=dev, generator=g, requires_grad=True)
        y = torch.randint(0, N, (B,), device=dev, generator=g)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        loss, log = crit(ev, ea, y)
        loss.backward()
        s1.record()
        torch.cuda.synchronize()
        if it >= 5:
            times.append(s0.elapsed_time(s1))
    times.sort()
    step_ms = times[len(times) // 2]
    bytes_ = 2.0 * B * (1 + 32 + 1024) * 512
    r = {"bank_rows": N, "mining_ms": mine_ms, "mining_TFLOPs": flops / (mine_ms * 1e-3) / 1e12, "mining_flops": flops,
         "constructor_s_incl_bank_init_and_first_mining": ctor_s,
         "criterion_step_ms_fwd_bwd_update": step_ms, "criterion_algorithmic_MB": bytes_ / 1e6, "criterion_GB/s": bytes_ / (step_ms * 1e-3) / 1e9,
         "loss": float(loss), "keys": sorted(log)}
    print(json.dumps(r))
    if a.out:
        json.dump(r, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
