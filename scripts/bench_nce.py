#!/usr/bin/env python
"""BASELINE.json config 5: NCE negative-count sweep K in {256, 1024, 4096, 16384} at batch 64 on one B200 -- the fused
gather + score + NCE + gradient kernel alone (avid_nce_forward_backward), reported as algorithmic GB/s
(2 banks x B x (K+1) x 512 B per call) against the measured HBM copy peak.

    python scripts/bench_nce.py [--banks 240000 2000000] [--iters 30] [--out profiles/r1_nce_sweep.json]

Negatives are drawn inside the kernel (Philox, a fresh offset per call) so no index tensor is read; the 2 M-row banks
(2 x 1.02 GB) do not fit the 126 MB L2, the 240 k banks (2 x 123 MB) partially do -- both are reported, and an L2 flush
(256 MB write) runs between timed calls.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--banks", type=int, nargs="+", default=[240000, 2000000])
    ap.add_argument("--negatives", type=int, nargs="+", default=[256, 1024, 4096, 16384])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--out", default=None)
    ap.add_argument("--flush", choices=["write", "read", "none"], default="write",
                    help="between timed calls: write 256 MB (L2 left full of DIRTY lines: the kernel's first misses each evict one), "
                         "read 256 MB (L2 left full of clean lines), or nothing (only meaningful with banks much larger than L2)")
    ap.add_argument("--group", type=int, default=1, help="launches per timed event pair (each with fresh y and a fresh Philox offset)")
    a = ap.parse_args()
    from avid_cma_b200 import ops
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    peak = 6533.0
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
    sink = torch.zeros((), dtype=torch.int64, device=dev)
    B = a.batch
    results = []
    for N in a.banks:
        g = torch.Generator(device=dev).manual_seed(N)
        bank_v = ops.rows_l2_normalize_(torch.randn(N, 128, device=dev, generator=g))
        bank_a = ops.rows_l2_normalize_(torch.randn(N, 128, device=dev, generator=g))
        emb_v, emb_a = torch.randn(B, 128, device=dev, generator=g), torch.randn(B, 128, device=dev, generator=g)
        Z = torch.full((), 2.0, device=dev)
        for K in a.negatives:
            keys = [(0, 1, 0, K, 0.5), (1, 0, 0, K, 0.5)]                  # Cross-N{K}: v2a, a2v
            ws = ops.nce_workspace(B, K, 0, len(keys), dev)
            out = torch.empty(1 + len(keys) + 2 * B * 128, device=dev)
            lt, lk = out[0:1], out[1:1 + len(keys)]
            gv, ga = out[1 + len(keys):1 + len(keys) + B * 128].view(B, 128), out[1 + len(keys) + B * 128:].view(B, 128)
            times = []
            for it in range(a.warmup + a.iters):
                calls = []
                for gi in range(a.group):
                    y = torch.randint(0, N, (B,), device=dev, generator=g)
                    calls.append(ops.make_nce_args(emb_v, emb_a, y, bank_v, bank_a, keys, K, Z, seed=1234, offset=(it * a.group + gi) * B * K,
                                                   loss_keys=lk, loss_total=lt, grad_v=gv, grad_a=ga))
                if a.flush == "write":
                    flush.fill_(it & 0xFF)
                elif a.flush == "read":
                    sink.copy_(flush.view(torch.int64).sum())
                else:
                    torch.cuda._sleep(200000)       # keeps the GPU busy while the host enqueues: no launch latency inside the event pair
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for args in calls:
                    ops.nce_forward_backward(args, ws)
                e1.record()
                torch.cuda.synchronize()
                if it >= a.warmup:
                    times.append(e0.elapsed_time(e1) / a.group)
            times.sort()
            ms = times[len(times) // 2]
            bytes_ = 2.0 * B * (K + 1) * 512
            r = {"bank_rows": N, "K": K, "batch": B, "ms_median": ms, "ms_min": times[0], "algorithmic_MB": bytes_ / 1e6,
                 "GB/s": bytes_ / (ms * 1e-3) / 1e9, "frac_of_measured_hbm": bytes_ / (ms * 1e-3) / 1e9 / peak, "hbm_peak_GB/s": peak,
                 "flush": a.flush, "launches_per_event_pair": a.group, "loss": float(lt)}
            results.append(r)
            print(json.dumps(r), flush=True)
        del bank_v, bank_a
    if a.out:
        with open(a.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
