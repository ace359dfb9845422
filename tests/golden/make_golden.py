#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported read-only from
/root/reference) on the deterministic synthetic inputs of oracle/synth.py.

Runs only in the build container (the GPU box has no /root/reference); the .npz outputs are
committed.  Usage:  python tests/golden/make_golden.py

The reference hard-codes `.cuda(device)` (avid.py:93,96,179; avid_cma.py:224,294); on this
GPU-less host `.cuda` is shimmed to a no-op before the import, no reference file is touched.
CMASampler.sample() needs visible GPUs to spawn workers (avid_cma.py:100-123); it is replaced
by an in-process loop that drives the reference's own sample_instance()/sample_gather() with
list-backed queues.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(1, "/root/reference")

if not torch.cuda.is_available():
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self

import criterions  # noqa: E402  (reference)
import models      # noqa: E402  (reference)
from criterions import avid_cma as ref_cma  # noqa: E402
from oracle import synth  # noqa: E402

torch.set_num_threads(8)


class _ListQueue:
    def __init__(self, items=()):
        self.items = list(items)

    def get(self):
        return self.items.pop(0)

    def put(self, x):
        self.items.append(x)


def _inprocess_sample(self):
    n = self.video_mem.shape[0]
    jobs = [list(range(i, min(i + 16, n))) for i in range(0, n, 16)] + [None]
    q_job, q_data = _ListQueue(jobs), _ListQueue()
    self.sample_instance(0, q_job, q_data)
    return self.sample_gather(q_data, workers=1)


ref_cma.CMASampler.sample = _inprocess_sample


def np_(t):
    return t.detach().cpu().numpy()


def criterion_case(tag, N, B, K, seed, xw=(1.0, 0.0), momentum=0.5, steps=2):
    """AVID criterion alone: `steps` consecutive forward/backward calls (Z frozen after the first)."""
    crit = criterions.AVID(num_data=N, embedding_dim=128, num_negatives=K, momentum=momentum,
                           xModal_coeff=xw[0], wModal_coeff=xw[1], device=0)
    bv, ba = synth.bank(N, seed=seed, tag="bank_v"), synth.bank(N, seed=seed, tag="bank_a")
    crit.nce_average.view1_mem.copy_(bv)
    crit.nce_average.view2_mem.copy_(ba)
    out = {"N": N, "B": B, "K": K, "seed": seed, "xModal": xw[0], "wModal": xw[1],
           "momentum": np.asarray(momentum, dtype=np.float64), "steps": steps}
    for s in range(steps):
        ev, ea = synth.embeddings(B, seed=seed + 100 * s)
        y = synth.instance_ids(B, N, seed=seed + 100 * s)
        idx = synth.negatives(y, K, N, seed=seed + 100 * s)
        crit.nce_average.sample_negatives = lambda y_, K_, idx=idx: idx
        ev.requires_grad_(True)
        ea.requires_grad_(True)
        loss, log = crit(ev, ea, y)
        loss.backward()
        out[f"s{s}_total"] = np_(loss)
        for k, v in log.items():
            out[f"s{s}_{k}"] = np.asarray(float(v))
        out[f"s{s}_grad_v"] = np_(ev.grad)
        out[f"s{s}_grad_a"] = np_(ea.grad)
        out[f"s{s}_Z"] = np.asarray(float(crit.criterion.avg_exp_score))
        out[f"s{s}_rows_v"] = np_(crit.nce_average.view1_mem[y])
        out[f"s{s}_rows_a"] = np_(crit.nce_average.view2_mem[y])
    np.savez_compressed(os.path.join(HERE, f"criterion_{tag}.npz"), **out)
    print("criterion", tag, {k: v for k, v in out.items() if k.endswith("total") or k.endswith("_Z")})


def cma_case(tag, N, B, K, Kw, pos_k, mode, seed):
    """AVID_CMA: positive mining (reference sample_instance), remapped negatives, 4-key loss."""
    torch.manual_seed(seed)
    crit = criterions.AVID_CMA(num_data=N, embedding_dim=128, num_negatives=K, num_negatives_within=Kw,
                               momentum=0.5, xModalInstCoeff=1., wModalInstCoeff=0., xModalPosCoeff=0.,
                               wModalPosCoeff=1., sampling_args={"type": mode, "pos_k": pos_k}, device=0)
    # banks with planted near-duplicates so the positive sets are non-trivial
    bv, ba = synth.bank(N, seed=seed, tag="bank_v"), synth.bank(N, seed=seed, tag="bank_a")
    crit.nce_average.view1_mem.copy_(bv)
    crit.nce_average.view2_mem.copy_(ba)
    crit.nce_average.find_correspondences()
    pos = crit.nce_average.positive_set.clone()
    out = {"N": N, "B": B, "K": K, "Kw": -1 if Kw is None else Kw, "pos_k": pos_k, "seed": seed, "positive_set": np_(pos)}
    ev, ea = synth.embeddings(B, seed=seed)
    y = synth.instance_ids(B, N, seed=seed)
    raw = synth.raw_negatives(B, K, N - pos_k, seed=seed)
    crit.nce_average.multinomial.draw = lambda n, raw=raw: raw.reshape(-1)
    _, neg = crit.nce_average.memory_sampling(y)
    out["neg_idx"] = np_(neg)
    ev.requires_grad_(True)
    ea.requires_grad_(True)
    loss, log = crit(ev, ea, y)
    loss.backward()
    out["total"] = np_(loss)
    for k, v in log.items():
        out[k] = np.asarray(float(v))
    out["grad_v"], out["grad_a"] = np_(ev.grad), np_(ea.grad)
    out["Z"] = np.asarray(float(crit.criterion.avg_exp_score))
    out["rows_v"] = np_(crit.nce_average.view1_mem[y])
    out["rows_a"] = np_(crit.nce_average.view2_mem[y])
    np.savez_compressed(os.path.join(HERE, f"cma_{tag}.npz"), **out)
    print("cma", tag, float(loss), {k: float(v) for k, v in log.items()})


def cma_mining_case(tag, N, pos_k, mode, seed):
    """Positive mining only (reference CMASampler.sample_instance, avid_cma.py:42-73) for the single-modality sampling types."""
    torch.manual_seed(seed)
    crit = criterions.AVID_CMA(num_data=N, embedding_dim=128, num_negatives=16, num_negatives_within=8, momentum=0.5,
                               sampling_args={"type": mode, "pos_k": pos_k}, device=0)
    crit.nce_average.view1_mem.copy_(synth.bank(N, seed=seed, tag="bank_v"))
    crit.nce_average.view2_mem.copy_(synth.bank(N, seed=seed, tag="bank_a"))
    crit.nce_average.find_correspondences()
    np.savez_compressed(os.path.join(HERE, f"cma_mining_{tag}.npz"), N=N, pos_k=pos_k, seed=seed,
                        positive_set=np_(crit.nce_average.positive_set))
    print("cma mining", tag, tuple(crit.nce_average.positive_set.shape))


def step_case(tag="config1", B=4, N=64, K=1024, size=112, spec=(100, 129), seed=0):
    """BASELINE config 1: full forward + AVID criterion + backward through both towers."""
    model = models.av_wrapper("R2Plus1D", {"depth": 18}, "Conv2D", {"depth": 10}, proj_dim=[512, 512, 128])
    sd = synth.fill_state_dict(model.state_dict(), seed=seed)
    model.load_state_dict(sd)
    model.train()
    crit = criterions.AVID(num_data=N, embedding_dim=128, num_negatives=K, momentum=0.5,
                           xModal_coeff=1., wModal_coeff=0., device=0)
    crit.nce_average.view1_mem.copy_(synth.bank(N, seed=seed, tag="bank_v"))
    crit.nce_average.view2_mem.copy_(synth.bank(N, seed=seed, tag="bank_a"))
    video, audio = synth.clips(B, 8, size, seed), synth.spectrograms(B, spec[0], spec[1], seed)
    y = torch.tensor([1, 17, 33, 60][:B]) if N == 64 else synth.instance_ids(B, N, seed)
    idx = synth.negatives(y, K, N, seed)
    crit.nce_average.sample_negatives = lambda y_, K_: idx
    ve, ae = model(video, audio)
    loss, log = crit(ve, ae, y)
    loss.backward()
    out = {"B": B, "N": N, "K": K, "size": size, "spec": np.asarray(spec), "seed": seed, "y": np_(y),
           "video_emb": np_(ve), "audio_emb": np_(ae), "total": np_(loss),
           "Z": np.asarray(float(crit.criterion.avg_exp_score)),
           "rows_v": np_(crit.nce_average.view1_mem[y]), "rows_a": np_(crit.nce_average.view2_mem[y])}
    for k, v in log.items():
        out[k] = np.asarray(float(v))
    # gradients: L2 norm of every parameter gradient + the full gradient of a few small tensors
    names, norms = [], []
    for n, p in model.named_parameters():
        names.append(n)
        norms.append(float(p.grad.double().norm()))
    out["grad_names"] = np.asarray(names)
    out["grad_norms"] = np.asarray(norms)
    for n in ("video_model.conv1.1.weight", "video_model.conv2x.0.spt_bn1.bias", "video_model.conv5x.1.out_bn.weight",
              "audio_model.conv1.1.bias", "audio_model.block4.bn2.weight", "video_proj.projection.4.bias",
              "audio_proj.projection.0.bias", "video_model.conv3x.0.res_conv.weight"):
        out["grad::" + n] = np_(dict(model.named_parameters())[n].grad)
    out["grad_slice::video_model.conv1.0.weight"] = np_(model.video_model.conv1[0].weight.grad[:4])
    out["grad_slice::audio_model.conv1.0.weight"] = np_(model.audio_model.conv1[0].weight.grad[:4])
    out["grad_slice::video_model.conv4x.1.tmp_conv2.weight"] = np_(model.video_model.conv4x[1].tmp_conv2.weight.grad[:2])
    # BN running statistics after the step
    msd = model.state_dict()
    for n in ("video_model.conv1.1", "video_model.conv5x.1.out_bn", "audio_model.block2.bn1"):
        out["rm::" + n] = np_(msd[n + ".running_mean"])
        out["rv::" + n] = np_(msd[n + ".running_var"])
    # fp64 re-computation with the oracle (pinned against the reference above): separates "reference fp32 rounding"
    # from "our error" -- with train-mode BN at batch 4 the early-layer gradients of two fp32 runs differ by ~4e-3.
    from oracle import criterion as oc, towers
    sd64 = synth.fill_state_dict(towers.state_dict_template(), seed=seed)
    for k in towers.param_keys(sd64):
        sd64[k] = sd64[k].double().requires_grad_(True)
    ve64, ae64 = towers.av_forward(video.double(), audio.double(), sd64, training=True)
    bank64 = [synth.bank(N, seed=seed, tag=t) for t in ("bank_v", "bank_a")]
    tot64, _, _ = oc.criterion_forward(ve64, ae64, y, bank64[0], bank64[1], idx, oc.avid_keys(K))
    tot64.backward()
    out["fp64::video_emb"], out["fp64::audio_emb"], out["fp64::total"] = np_(ve64), np_(ae64), np_(tot64)
    out["fp64::grad_norms"] = np.asarray([float(sd64[n].grad.norm()) for n in names])
    for k in list(out):
        if k.startswith("grad::"):
            out["fp64::" + k] = np_(sd64[k[6:]].grad)
    np.savez_compressed(os.path.join(HERE, f"step_{tag}.npz"), **out)
    print("step", tag, "loss", float(loss), "Z", float(out["Z"]))


if __name__ == "__main__":
    criterion_case("cross", N=512, B=8, K=64, seed=1)
    criterion_case("joint", N=300, B=5, K=33, seed=2, xw=(1.0, 1.0), momentum=[0.3, 0.8])
    criterion_case("cfg1", N=64, B=4, K=1024, seed=3, steps=1)
    cma_case("consensus", N=400, B=6, K=96, Kw=16, pos_k=8, mode="consensus", seed=4)
    cma_case("union", N=257, B=4, K=40, Kw=None, pos_k=5, mode="union", seed=5)
    cma_mining_case("video", N=333, pos_k=6, mode="video", seed=6)
    cma_mining_case("audio", N=190, pos_k=9, mode="audio", seed=7)
    step_case()
