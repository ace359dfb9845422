#!/usr/bin/env python
"""Headline benchmark: AVID training clips/sec (video+audio pair) on N B200s (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One step = the reference's run_phase body (main-avid.py:155-184): forward of both towers, AVID criterion
(fused gather + NCE + bank update), backward, Adam.  Workload = BASELINE.json configs[1]: Cross-N1024 AVID,
240k-entry memory bank, batch 64/GPU of 8x3x224x224 clips + 1x200x257 spectrograms, synthetic data, random-init
weights.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "AVID training clips/sec (video+audio pair)"
UNIT = "clips/s"
BANK_ROWS = 240000
NUM_NEG = 1024
FWD_GFLOP_PER_CLIP = 27.134       # SURVEY.md §8d (conv + linear, 2*MAC) at 8x3x224x224 + 1x200x257
STEP_GFLOP_PER_CLIP = 75.66


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_gpu"],
                    help="ours: this package; reference: the reference's CPU path (oracle port) on the host cores; torch_gpu: DIAGNOSTIC arm, the "
                         "same reference call sequence through stock torch (cuDNN / cuBLAS / ATen) on cuda:0, TF32 off and on")
    ap.add_argument("--batch", type=int, default=64, help="clips per GPU per step")
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--spec", type=int, nargs=2, default=[200, 257])
    ap.add_argument("--bank", type=int, default=BANK_ROWS)
    ap.add_argument("--negatives", type=int, default=NUM_NEG)
    ap.add_argument("--math", default=os.environ.get("AVID_MATH", "bf16x3"), choices=["fp32", "bf16x3", "bf16"])
    ap.add_argument("--cpu-sample-batch", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph-experiment", action="store_true",
                    help="diagnostic: also time K replays of ONE training step captured in a CUDA graph (negatives / Adam step count frozen: timing only)")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the stock-torch GPU leg (gpu_library_baseline) of the N=1 line")
    ap.add_argument("--no-subrecords", action="store_true", help="N > 1: skip the config-3 (2 M-row sharded bank) and config-4 (sharded CMA) sub-records")
    ap.add_argument("--sub-steps", type=int, default=8, help="timed steps of each N > 1 sub-record (after 3 warm-up steps)")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: skip the end-to-end timed region")
    ap.add_argument("--dump-launches", default=None, help="write the per-launch CUDA-event table of the timed region (mean over steps, launch order) to this file")
    ap.add_argument("--grad-sync", default="auto", choices=["auto", "fused", "ddp"],
                    help="N > 1: fused = gradients exchanged inside the optimizer step over NVLink peer memory (optim.ShardedAdam, default when "
                         "symmetric memory is available), ddp = the reference's DistributedDataParallel all-reduce + Adam")
    ap.add_argument("--bank-mode", default="auto", choices=["auto", "replicated", "sharded"],
                    help="memory-bank layout for N > 1: sharded = row-partitioned over the ranks (default), replicated = the reference's")
    return ap.parse_args()


def config_of(a, n_gpus):
    return {"workload": "Cross-N1024 AVID, 240k-entry memory bank (Kinetics-shape), batch=64/GPU 8x3x224x224 + 1x200x257",
            "global_batch": a.batch * n_gpus, "batch_per_gpu": a.batch, "clip": [3, a.frames, a.size, a.size], "spectrogram": [1] + list(a.spec),
            "bank_rows": a.bank, "num_negatives": a.negatives, "optimizer": "adam lr 2e-4 wd 1e-5",
            "parallelism": f"dp{n_gpus}", "grad_sync": getattr(a, "grad_sync_used", "none"), "bank_layout": "single" if n_gpus == 1 else ("replicated" if a.bank_mode == "replicated" else "row-sharded"), "math": a.math, "tower_streams": 2 if os.environ.get("AVID_TOWER_STREAMS", "1") == "1" else 1, "l2": "inputs larger than L2 (one batch of clips = %.0f MB)" % (a.batch * 3 * a.frames * a.size * a.size * 4 / 1e6)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference(a, steps, warmup, batch):
    """The reference's CPU path (oracle port: torch-CPU restatement of towers + criterion + Adam) on this box's host cores."""
    from oracle import synth
    from oracle.step import OracleTrainer
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tr = OracleTrainer(a.bank, num_negatives=a.negatives, seed=0)
    g = torch.Generator().manual_seed(0)
    video = torch.randn(batch, 3, a.frames, a.size, a.size, generator=g)
    audio = torch.randn(batch, 1, a.spec[0], a.spec[1], generator=g)
    for i in range(warmup):
        tr.step(video, audio, synth.instance_ids(batch, a.bank, seed=i))
    t0 = time.perf_counter()
    for i in range(steps):
        tr.step(video, audio, synth.instance_ids(batch, a.bank, seed=100 + i))
    dt = time.perf_counter() - t0
    return {"value": batch * steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{steps} full training steps (fwd + AVID criterion + bwd + Adam) of {batch} clips at the workload's clip/spectrogram shape, "
                      f"bank {a.bank}, K={a.negatives}, torch-CPU oracle port of the reference, {cores} threads, after {warmup} warm-up"}, dt / steps


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, sec = cpu_reference(a, a.steps, a.warmup, a.cpu_sample_batch)
    cfg = config_of(a, a.gpus)
    # this arm is ONE host process stepping a bounded sample of the workload: say so in ITS config (same clip / spectrogram shape,
    # bank, K and optimizer; clips/s on the CPU is batch-independent to a few per cent)
    cfg.update({"global_batch": a.cpu_sample_batch, "batch_per_gpu": a.cpu_sample_batch, "parallelism": "1 host process, %d threads" % cb["cores"],
                "bank_layout": "single", "math": "f32 (torch CPU)",
                "sample": "%d-clip steps of the workload (the GPU arm steps %d clips per GPU)" % (a.cpu_sample_batch, a.batch)})
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg, "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ stock-torch GPU arm (diagnostic)
def torch_gpu_baseline(a, steps, warmup, device):
    """the reference's training step through stock torch 2.11 on the same B200 -- cuDNN conv3d / conv2d,
    cuBLAS bmm / Linear, the ~150 small ATen launches of the criterion, host-drawn negatives copied each step, torch.optim.Adam --
    with cudnn.benchmark on (main-avid.py:121), TF32 off (fp32 truth) and on (torch's default for convolutions).  The oracle port
    is that call sequence as functional torch; here it only runs on `device`.  DIAGNOSTIC: not the driver's reference arm."""
    from oracle import synth
    from oracle.step import OracleTrainer
    out = {}
    B = a.batch
    g = torch.Generator().manual_seed(0)
    video = torch.randn(B, 3, a.frames, a.size, a.size, generator=g).to(device)
    audio = torch.randn(B, 1, a.spec[0], a.spec[1], generator=g).to(device)
    saved = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        torch.backends.cudnn.benchmark = True
        for tf32 in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            tr = OracleTrainer(a.bank, num_negatives=a.negatives, seed=0, device=device)
            ys = [synth.instance_ids(B, a.bank, seed=i).to(device) for i in range(warmup + steps)]
            for i in range(warmup):
                tr.step(video, audio, ys[i])
            torch.cuda.synchronize(device)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                tr.step(video, audio, ys[warmup + i])
            e1.record()
            torch.cuda.synchronize(device)
            ms = e0.elapsed_time(e1) / steps
            out["tf32_on" if tf32 else "fp32"] = {"clips_per_s": B / (ms * 1e-3), "ms_per_step": ms}
            del tr
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    out.update({"unit": UNIT, "steps": steps, "warmup": warmup, "batch": B,
                "what": "the reference's step (oracle port = its torch call sequence) on stock torch %s CUDA: cuDNN / cuBLAS / ATen, "
                        "cudnn.benchmark=True, inputs resident, loss read back every step" % torch.__version__})
    return out


def input_side(a, dev):
    """SURVEY 8f-3 sub-record: the input-side steps on the GPU with the reference's CPU path (Pillow through torchvision's wrappers, oracle
    port of the librosa spectrogram) timed beside them on a bounded sample.  Video: one batch of 64 decoded clips (8 x 256 x 340 uint8,
    the Kinetics short-side-256 frames) -> (3, 8, 224, 224) float32 with VideoPrep_MSC_CJ's crop / flip / jitter / normalise."""
    import random
    import time
    import numpy as np
    from avid_cma_b200.datasets.gpu_preprocessing import VideoPrep_MSC_CJ
    from oracle import video as OV
    B, T, H, W = a.batch, a.frames, 256, 340
    g = np.random.default_rng(0)
    host = g.integers(0, 256, (B, T, H, W, 3), dtype=np.uint8)
    clips = torch.from_numpy(host).to(dev)
    prep = VideoPrep_MSC_CJ(crop=(a.size, a.size), num_frames=T)
    random.seed(0)
    params = [prep.draw(W, H) for _ in range(B)]
    for _ in range(3):
        prep.apply_batch(list(clips), params)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    e0.record()
    for _ in range(iters):
        outs = prep.apply_batch(list(clips), params)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / iters
    run, _ = prep.plan_batch(list(clips), params)      # the kernels alone: host-side planning (ctypes structs, buffers) outside the timed region
    run()
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize(dev)
    ms_k = e0.elapsed_time(e1) / iters
    # algorithmic bytes per clip: the crop is read once (uint8), the result written once (float32)
    alg = sum(T * q['crop'][2] * q['crop'][3] * 3 for q in params) + B * 3 * T * a.size * a.size * 4
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    n_cpu = 4
    t0 = time.perf_counter()
    for k in range(n_cpu):
        ref = OV.video_prep_pil(host[k], params[k], crop=(a.size, a.size))
    cpu_s = (time.perf_counter() - t0) / n_cpu
    exact = bool(np.array_equal(outs[n_cpu - 1].cpu().numpy(), ref))
    return {"video": {"what": "VideoPrep_MSC_CJ (crop + Pillow-exact bilinear resize + flip + colour jitter + normalise), %d clips of %dx%dx%d uint8 -> %dx%d float32, one avid_video_prep_batch call" % (B, T, H, W, a.size, a.size),
                      "clips_per_s": B / (ms * 1e-3), "ms_per_batch": ms, "ms_per_batch_kernels_only": ms_k, "algorithmic_bytes_per_batch": alg,
                      "roofline": {"bound": "hbm", "achieved": alg / (ms_k * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": alg / (ms_k * 1e-3) / 1e9 / hbm,
                                   "note": "kernels only (4 launches per 16 clips); the public call adds the per-clip host planning"},
                      "cpu_baseline": {"value": 1.0 / cpu_s, "unit": "clips/s", "cores": 1, "kind": "reference",
                                       "sample": "%d clips through Pillow %s (the library the reference's transforms execute), one thread" % (n_cpu, __import__("PIL").__version__)},
                      "bit_exact_vs_pillow": exact}}


def run_torch_gpu(a):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    res = torch_gpu_baseline(a, a.steps, a.warmup, dev)
    line = {"impl": "torch_gpu", "metric": METRIC, "value": res["fp32"]["clips_per_s"], "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": res["fp32"]["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (TF32 off)",
            "data": "synthetic", "config": config_of(a, 1), "gpu_library_baseline": res}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        mhz = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons, "samples": len(mhz)}


# ------------------------------------------------------------------------------------------------ our arm
def _max_over_ranks(vals, dev, world):
    import torch.distributed as dist
    t = torch.tensor(vals, device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def parity_check(net, video, audio, a, world, rank, local, dev):
    """N > 1, before anything is timed: ONE gathered batch through the row-sharded criterion (the layout this run times) and through
    the replicated one (the reference's layout, avid.py:103-129), same Philox stream, same banks (rows are a function of the shared
    seed).  Loss, d loss / d embeddings and the updated bank rows must agree to fp32 summation order; the run FAILS above 1e-4."""
    import torch.distributed as dist
    from avid_cma_b200.criterions import AVID
    crits = {}
    for mode in ("sharded", "replicated"):
        os.environ["AVID_SHARD_BANK"] = "1" if mode == "sharded" else "0"
        crits[mode] = AVID(num_data=a.bank, embedding_dim=128, num_negatives=a.negatives, momentum=0.5, xModal_coeff=1., wModal_coeff=0., device=local)
    assert crits["sharded"].nce_average.sharded and not crits["replicated"].nce_average.sharded
    with torch.no_grad():
        ve, ae = net(video, audio)
    y = torch.randperm(a.bank, generator=torch.Generator().manual_seed(99))[rank * a.batch:(rank + 1) * a.batch].to(dev)
    res = {}
    for mode, crit in crits.items():
        ev, ea = ve.detach().clone().requires_grad_(True), ae.detach().clone().requires_grad_(True)
        loss, _ = crit(ev, ea, y)
        loss.backward()
        res[mode] = (loss.detach().double(), ev.grad.double(), ea.grad.double())
    ls, lr = res["sharded"][0], res["replicated"][0]
    loss_rel = float((ls - lr).abs() / lr.abs())
    grad_rel = max(float((res["sharded"][i] - res["replicated"][i]).norm() / res["replicated"][i].norm()) for i in (1, 2))
    bs, br = crits["sharded"].nce_average, crits["replicated"].nce_average
    bank_rel = max(float((getattr(bs, n).double() - getattr(br, n)[bs.row_begin:bs.row_end].double()).norm() /
                         getattr(br, n)[bs.row_begin:bs.row_end].double().norm()) for n in ("view1_mem", "view2_mem"))
    z_rel = abs(float(crits["sharded"].criterion.avg_exp_score) - float(crits["replicated"].criterion.avg_exp_score)) / float(crits["replicated"].criterion.avg_exp_score)
    loss_rel, grad_rel, bank_rel, z_rel = _max_over_ranks([loss_rel, grad_rel, bank_rel, z_rel], dev, world)
    out = {"loss_rel": loss_rel, "grad_rel": grad_rel, "bank_rel": bank_rel, "z_rel": z_rel, "tolerance": 1e-4,
           "what": "step 0 of one gathered batch: row-sharded criterion vs replicated criterion (max over ranks), in-kernel Philox negatives"}
    del crits
    torch.cuda.empty_cache()
    if not max(loss_rel, grad_rel, bank_rel, z_rel) <= 1e-4:
        raise RuntimeError("sharded criterion != replicated criterion: %s" % json.dumps(out))
    os.environ["AVID_SHARD_BANK"] = "0" if a.bank_mode == "replicated" else "1"
    return out


def sub_record(name, crit, net, opt, resident, a, world, rank, dev, rows):
    """3 warm-up + a.sub_steps timed training steps of the SAME towers with another criterion (BASELINE configs 3 and 4)."""
    import torch.distributed as dist
    from avid_cma_b200 import ops
    B = a.batch
    perm = torch.randperm(rows, generator=torch.Generator().manual_seed(11))
    ys = [perm[(i * world + rank) * B % (rows - B):][:B].contiguous().to(dev) for i in range(3 + a.sub_steps + 3)]

    marks = []          # (forward+criterion start, criterion start, criterion end, step end) CUDA events of the timed steps

    def step(i, timed=False):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timed else None
        if timed:
            ev[0].record()
        ve, ae = net(*resident[i % len(resident)])
        if timed:
            ev[1].record()
        loss, _ = crit(ve, ae, ys[i])
        if timed:
            ev[2].record()
        opt.zero_grad()
        loss.backward()
        opt.step()
        if timed:
            ev[3].record()
            marks.append(ev)
        return loss

    for i in range(3):
        step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3, 3 + a.sub_steps):
        loss = step(i, timed=True)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms, = _max_over_ranks([e0.elapsed_time(e1)], dev, world)
    crit_ms = sum(m[1].elapsed_time(m[2]) for m in marks) / len(marks)
    fwd_ms = sum(m[0].elapsed_time(m[1]) for m in marks) / len(marks)
    bwd_ms = sum(m[2].elapsed_time(m[3]) for m in marks) / len(marks)
    # three more steps with a CUDA-event pair around the fused criterion launch (not part of the timed region: per-launch events
    # switch the towers' CUDA graphs off)
    ops.profile_begin()
    for i in range(3 + a.sub_steps, 3 + a.sub_steps + 3):
        step(i)
    prof = ops.profile_end()
    nce = [(w, d) for n_, w, d in prof if n_ == "nce_fused"]
    rec = {"clips_per_s": B * world * a.sub_steps / (ms * 1e-3), "ms_per_step": ms / a.sub_steps, "steps": a.sub_steps, "warmup": 3,
           "last_loss": float(loss.detach()), "ms_towers_forward": fwd_ms, "ms_criterion": crit_ms, "ms_backward_optimizer": bwd_ms}
    if nce:
        rec["nce_GBps_per_gpu"] = sum(w for w, _ in nce) / (sum(d for _, d in nce) * 1e-3) / 1e9
        rec["nce_us_per_launch"] = 1e3 * sum(d for _, d in nce) / len(nce)
        rec["nce_algorithmic_bytes_per_launch_per_gpu"] = nce[0][0]
    return rec


def sub_records(net, opt, resident, a, world, rank, local, dev):
    """BASELINE.json configs 3 (Audioset-shape 2 M-row bank, row-sharded) and 4 (AVID+CMA, 240 k bank, top-32 consensus mining
    sharded over the ranks) on the towers this run has just timed."""
    import torch.distributed as dist
    from avid_cma_b200.criterions import AVID, AVID_CMA
    os.environ["AVID_SHARD_BANK"] = "1"
    out = {}

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- config 3: Cross-N1024 AVID, 2 M-entry sharded bank (configs/main/avid/audioset/Cross-N1024.yaml: num_data 1 784 108)
    N3 = 2000000
    sync()
    t0 = time.perf_counter()
    crit3 = AVID(num_data=N3, embedding_dim=128, num_negatives=a.negatives, momentum=0.5, xModal_coeff=1., wModal_coeff=0., device=local)
    sync()
    init_s, = _max_over_ranks([time.perf_counter() - t0], dev, world)
    rec = sub_record("config3", crit3, net, opt, resident, a, world, rank, dev, N3)
    rec.update({"workload": "Cross-N1024 AVID, 2M-entry sharded memory bank (Audioset-shape), batch=64/GPU, NCCL packed all-gather + reduce-scatter",
                "bank_rows": N3, "rows_per_gpu": crit3.nce_average.rows_per_rank, "bank_init_s": init_s,
                "bank_init": "per-rank seeded rows (avid_bank_init), no (N,128) broadcast"})
    out["config3"] = rec
    del crit3
    torch.cuda.empty_cache()

    # ---- config 4: InstX-N1024-PosW-N64-Top32 AVID+CMA, 240 k bank (configs/main/avid-cma/kinetics/InstX-N1024-PosW-N64-Top32.yaml:47-62)
    N4 = a.bank
    sync()
    t0 = time.perf_counter()
    crit4 = AVID_CMA(num_data=N4, embedding_dim=128, num_negatives=a.negatives, num_negatives_within=64, momentum=0.5,
                     xModalInstCoeff=1., wModalInstCoeff=0., xModalPosCoeff=0., wModalPosCoeff=1.,
                     sampling_args={"type": "consensus", "pos_k": 32}, resample_freq=-1, device=local)
    sync()
    build_s, = _max_over_ranks([time.perf_counter() - t0], dev, world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    crit4.nce_average.find_correspondences()
    e1.record()
    sync()
    mining_ms, = _max_over_ranks([e0.elapsed_time(e1)], dev, world)
    rec = sub_record("config4", crit4, net, opt, resident, a, world, rank, dev, N4)
    rec.update({"workload": "InstX-N1024-PosW-N64-Top32 AVID+CMA, 240k bank, CMA top-32 consensus positive expansion, mining sharded over the ranks",
                "bank_rows": N4, "mining_ms": mining_ms, "mining_TFLOPs": 2 * 2.0 * N4 * N4 * 128 / (mining_ms * 1e-3) / 1e12,
                "criterion_build_s": build_s})
    out["config4"] = rec
    del crit4
    torch.cuda.empty_cache()
    os.environ["AVID_SHARD_BANK"] = "0" if a.bank_mode == "replicated" else "1"
    return out


def run_ours(a):
    import torch.distributed as dist
    from avid_cma_b200 import models, ops, optim
    from avid_cma_b200.criterions import AVID
    from avid_cma_b200.models import _tower

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    os.environ["AVID_MATH"] = a.math

    torch.manual_seed(0)
    model = models.av_wrapper('R2Plus1D', {'depth': 18}, 'Conv2D', {'depth': 10}, proj_dim=[512, 512, 128]).to(dev).train()
    os.environ["AVID_SHARD_BANK"] = "0" if a.bank_mode == "replicated" else "1"
    crit = AVID(num_data=a.bank, embedding_dim=model.out_dim, num_negatives=a.negatives, momentum=0.5, xModal_coeff=1., wModal_coeff=0., device=local)
    # broadcast_buffers=False: the per-forward broadcast of rank 0's 126 BatchNorm buffers only matters for what a checkpoint holds,
    # and rank 0 writes the checkpoint either way (utils/main_utils.py CheckpointManager)
    ddp = world > 1 and os.environ.get("AVID_BENCH_NO_DDP", "0") != "1"       # AVID_BENCH_NO_DDP=1: diagnostic only (no gradient sync)
    fused = ddp and a.grad_sync != "ddp" and optim.ShardedAdam.available()
    if a.grad_sync == "fused" and world > 1 and not fused:
        raise RuntimeError("--grad-sync fused needs torch symmetric memory (all ranks on one node, NCCL backend)")
    a.grad_sync_used = "none" if world == 1 else ("fused reduce-scatter + Adam + all-gather over NVLink peer memory (ShardedAdam)" if fused else
                                                  ("DistributedDataParallel all-reduce + Adam" if ddp else "off (diagnostic)"))
    if fused:
        net = optim.LocalGradients(model)
        opt = optim.ShardedAdam(model.parameters(), lr=2e-4, weight_decay=1e-5)
    else:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], broadcast_buffers=False, gradient_as_bucket_view=True) if ddp else model
        opt = optim.Adam(model.parameters(), lr=2e-4, weight_decay=1e-5)

    B = a.batch
    g = torch.Generator().manual_seed(1234 + rank)
    nbuf = 2
    host = [(torch.randn(B, 3, a.frames, a.size, a.size, generator=g).pin_memory(), torch.randn(B, 1, a.spec[0], a.spec[1], generator=g).pin_memory())
            for _ in range(nbuf)]
    resident = [(v.to(dev), s.to(dev)) for v, s in host]
    perm = torch.randperm(a.bank, generator=torch.Generator().manual_seed(7))
    total_steps = 3 * (a.warmup + a.steps) + 16
    ys_host = [perm[(i * world + rank) * B % (a.bank - B):][:B].contiguous().pin_memory() for i in range(total_steps)]
    ys_dev = [y.to(dev) for y in ys_host]

    def step(video, audio, y):
        ve, ae = net(video, audio)
        loss, _ = crit(ve, ae, y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    parity = None
    if world > 1 and a.bank_mode != "replicated":
        parity = parity_check(net, resident[0][0], resident[0][1], a, world, rank, local, dev)

    it = 0
    for _ in range(a.warmup):
        step(*resident[it % nbuf], ys_dev[it]); it += 1
    # ---- timed region 1: inputs resident in HBM (the headline `value`) ----
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ops.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nvtx_id = torch.cuda.nvtx.range_start("avid_timed")   # ncu --nvtx --nvtx-include "avid_timed": steady-state steps (a start/end range
    #                                                       spans threads: the backward kernels are launched by autograd's worker thread)
    e0.record()
    for _ in range(a.steps):
        step(*resident[it % nbuf], ys_dev[it]); it += 1
    e1.record()
    barrier()
    torch.cuda.nvtx.range_end(nvtx_id)
    ms = e0.elapsed_time(e1)
    launches = ops.launch_count()
    clocks = sampler.summary()

    # ---- timed region 1b: the same steps with a CUDA-event pair around every tensor-core / criterion launch (roofline).  The two
    #      towers normally run on two streams (models/av_wrapper.py); per-launch durations are only meaningful when the launches do
    #      not overlap, so this region runs them on one stream -- which is also why it is not the region `value` comes from.
    prof_steps = min(a.steps, 10)
    streams_env = os.environ.get("AVID_TOWER_STREAMS")
    os.environ["AVID_TOWER_STREAMS"] = "0"
    step(*resident[it % nbuf], ys_dev[it]); it += 1
    barrier()
    ops.profile_begin()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(prof_steps):
        step(*resident[it % nbuf], ys_dev[it]); it += 1
    p1.record()
    barrier()
    ms_prof = p0.elapsed_time(p1)
    prof = ops.profile_end()
    if streams_env is None:
        os.environ.pop("AVID_TOWER_STREAMS", None)
    else:
        os.environ["AVID_TOWER_STREAMS"] = streams_env

    graph_ms, graph_err = None, None
    if a.graph_experiment and world == 1:
        # the whole step in ONE graph: the towers' own graphs cannot be replayed inside a capture, so they run as plain launches here
        graph_env = os.environ.get("AVID_CUDA_GRAPH")
        os.environ["AVID_CUDA_GRAPH"] = "0"
        try:
            sv, sa, sy = resident[0][0].clone(), resident[0][1].clone(), ys_dev[0].clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    step(sv, sa, sy)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step(sv, sa, sy)
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            eg0, eg1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            eg0.record()
            for _ in range(a.steps):
                g.replay()
            eg1.record()
            torch.cuda.synchronize()
            graph_ms = eg0.elapsed_time(eg1) / a.steps
            del g
        except Exception as e:   # noqa: BLE001 -- diagnostic only: anything in the step that cannot be captured (host-side checks) ends it
            graph_err = repr(e)[:200]
            try:
                torch.cuda.synchronize()
            except Exception:   # noqa: BLE001
                pass
        if graph_env is None:
            os.environ.pop("AVID_CUDA_GRAPH", None)
        else:
            os.environ["AVID_CUDA_GRAPH"] = graph_env

    # ---- timed region 2: end to end, pinned host buffers -> H2D each step (side stream, one step ahead, like a prefetching
    #      loader with non_blocking copies: main-avid.py:161-163), loss.item() each step ----
    e2e_steps = 0 if a.skip_e2e else a.steps
    copy_stream = torch.cuda.Stream()

    def prefetch(k):
        v, s = host[k % nbuf]
        with torch.cuda.stream(copy_stream):
            dv, ds, dy = v.to(dev, non_blocking=True), s.to(dev, non_blocking=True), ys_host[k].to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return dv, ds, dy, ev

    def e2e_loop(n):
        nonlocal it
        last = None
        nxt = prefetch(it) if n else None
        for i in range(n):
            dv, ds, dy, ev = nxt
            nxt = prefetch(it + 1) if i + 1 < n else None
            torch.cuda.current_stream().wait_event(ev)
            for t in (dv, ds, dy):
                t.record_stream(torch.cuda.current_stream())
            last = step(dv, ds, dy).item()
            it += 1
        return last

    e2e_loop(min(2, a.warmup) if e2e_steps else 0)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    last_loss = e2e_loop(e2e_steps)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3) if e2e_steps else float("nan")
    h2d = sum(t.numel() * t.element_size() for t in host[0]) + ys_host[0].numel() * 8

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    subs = None
    if world > 1 and not a.no_subrecords and a.bank_mode != "replicated":
        subs = sub_records(net, opt, resident, a, world, rank, local, dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    clips = B * world * a.steps
    if a.dump_launches and len(prof) % prof_steps == 0:
        per = len(prof) // prof_steps
        with open(a.dump_launches, "w") as f:
            for i in range(per):
                durs = [prof[s_ * per + i][2] for s_ in range(prof_steps)]
                name, work = prof[i][0], prof[i][1]
                mean = sum(durs) / len(durs)
                f.write("%3d %-18s %9.1f us  work %10.3e  rate %8.1f (TFLOP/s or GB/s)\n" % (i, name, mean * 1e3, work, work / (mean * 1e-3) / (1e9 if name.startswith("nce") else 1e12)))
    # roofline of the dominant kernel, from CUDA events recorded around every launch of the timed region.  Families are the
    # instrumented entry points of ops.py; forward and input gradient of a kernel instantiation (conv_tc_kernel<128>, <64>,
    # conv_pair_kernel) are merged.
    fam = {}
    for name, work, dur in prof:
        f = fam.setdefault(name, [0.0, 0.0, 0])
        f[0] += work; f[1] += dur; f[2] += 1
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tensor_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    roofline, families = None, {}
    for name, (work, dur, cnt) in fam.items():
        rate = work / (dur * 1e-3) if dur > 0 else 0.0
        families[name] = {"launches": cnt, "ms_per_step": dur / prof_steps, "share_of_step": dur / ms_prof}
        families[name]["GB/s" if name.startswith("nce") else "TFLOP/s"] = rate / (1e9 if name.startswith("nce") else 1e12)
    kernels = {}
    for name, (work, dur, cnt) in fam.items():
        kname = {"conv_forward_tc128": "conv_tc_kernel<128>", "conv_dgrad_tc128": "conv_tc_kernel<128>", "conv_forward_tc64": "conv_tc_kernel<64>",
                 "conv_dgrad_tc64": "conv_tc_kernel<64>", "conv_pair_forward": "conv_pair_kernel",
                 "conv_pair_dgrad": "conv_pair_kernel", "conv_wgrad_tc": "wgrad_tc_kernel",
                 "stem_forward_tc": "stem_forward_kernel", "stem_wgrad_tc": "stem_wgrad_kernel", "conv_forward": "conv_igemm_kernel",
                 "conv_dgrad": "conv_igemm_kernel", "conv_wgrad": "conv_wgrad_kernel"}.get(name)
        if kname:
            k = kernels.setdefault(kname, [0.0, 0.0, 0])
            k[0] += work; k[1] += dur; k[2] += cnt
    if kernels:
        top = max(kernels, key=lambda n: kernels[n][1])
        work, dur, cnt = kernels[top]
        ach = work / (dur * 1e-3) / 1e12
        mma_factor = 3 if a.math == "bf16x3" and top != "conv_igemm_kernel" else 1
        traffic, traffic_src = None, None
        try:   # DRAM bytes per launch of this kernel from the committed ncu pass of the same command (profiles/)
            prof_k = json.load(open(os.path.join(ROOT, "profiles", "r2_%s_kernels.json" % a.math)))["kernels"]
            rows = [r for r in prof_k if top in r["kernel"]]
            n_l = sum(r["launches_per_step"] for r in rows)
            traffic = sum((r["dram_read_MB_per_launch"] + r["dram_write_MB_per_launch"]) * 1e6 * r["launches_per_step"] for r in rows) / n_l
            traffic_src = "profiles/r2_%s_kernels.json (steady-state ncu launch list of this command: dram__bytes_read.sum + dram__bytes_write.sum, mean per launch)" % a.math
        except Exception:
            pass
        roofline = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": tensor_peak, "unit": "TFLOP/s", "frac": ach / tensor_peak,
                    "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step), of measured" if peaks else "fallback",
                    "launches_timed": cnt, "avg_launch_ms": dur / cnt, "share_of_step": dur / ms_prof,
                    "measured_in": "%d single-stream steps with a CUDA-event pair around each launch (%.2f ms per step); `value` comes from the "
                                   "region before it (towers on two streams, no per-launch events)" % (prof_steps, ms_prof / prof_steps),
                    "algorithmic_flops_per_launch": work / cnt,
                    "executed_mma_factor": mma_factor, "executed_frac": ach * mma_factor / tensor_peak,
                    "note": "achieved counts each product once (2*MAC of the convolution); bf16x3 issues 3 bf16 MMAs per product, so the "
                            "tensor pipe executes executed_mma_factor x that",
                    "families": families}
        if "nce_fused" in fam:
            w_, d_, c_ = fam["nce_fused"]
            hbm = peaks.get("hbm_gbs", 6650.0)
            roofline["nce"] = {"kernel": "nce_gather_kernel (one launch: gather + score + NCE + gradient + reduce)", "bound": "hbm", "achieved": w_ / (d_ * 1e-3) / 1e9,
                               "peak": hbm, "unit": "GB/s", "frac": w_ / (d_ * 1e-3) / 1e9 / hbm, "algorithmic_bytes_per_launch": w_ / c_,
                               "sweep": "profiles/r2_nce_sweep.json (2 M-row banks, K = 256 / 1024 / 4096 / 16384: 0.10 / 0.28 / 0.54 / 0.76 of measured HBM)"}

    line = {"metric": METRIC, "value": clips / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "bf16x3": "bf16x3 (3-term split, f32 accumulate)", "bf16": "bf16"}[a.math], "data": "synthetic",
            "config": config_of(a, world), "clocks": clocks,
            "e2e": {"value": clips / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / a.steps},
            "gpu_launches": launches, "roofline": roofline, "last_loss": last_loss,
            "step_tflops": clips * STEP_GFLOP_PER_CLIP * 1e9 / (ms * 1e-3) / 1e12 / world}
    if graph_err is not None:
        line["graph_experiment"] = {"error": graph_err}
    if graph_ms is not None:
        line["graph_experiment"] = {"ms_per_step": graph_ms, "clips_per_s": B / (graph_ms * 1e-3),
                                    "note": "one captured step replayed (frozen negatives / Adam step count): launch-gap diagnostic, not a bench value"}
    if parity is not None:
        line["parity_check"] = parity
    if subs is not None:
        line.update(subs)
    if world == 1 and not a.no_gpu_baseline:
        try:   # the reference's step through stock torch on this very GPU, measured in the same run (diagnostic)
            line["gpu_library_baseline"] = torch_gpu_baseline(a, 5, 3, dev)
            line["gpu_library_baseline"]["ours_over_fp32"] = line["value"] / line["gpu_library_baseline"]["fp32"]["clips_per_s"]
            line["gpu_library_baseline"]["ours_over_tf32"] = line["value"] / line["gpu_library_baseline"]["tf32_on"]["clips_per_s"]
        except Exception as e:   # noqa: BLE001 -- a diagnostic leg must not take the headline line down
            line["gpu_library_baseline"] = {"error": repr(e)[:300]}
    if world == 1 and not a.no_cpu_baseline:
        cb, _ = cpu_reference(a, 6, 1, a.cpu_sample_batch)
        line["cpu_baseline"] = cb
        try:   # the input-side steps, measured beside the step they feed (diagnostic sub-record)
            line["input_side"] = input_side(a, dev)
        except Exception as e:   # noqa: BLE001
            line["input_side"] = {"error": repr(e)[:300]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch_gpu":
        run_torch_gpu(args)
    else:
        run_ours(args)
