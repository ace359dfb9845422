"""CUDA-graph replay of the towers (models/_tower.py::GraphedTower) == eager launches: the same 6 training steps (fresh inputs every
step, Adam updating the parameters in place, BatchNorm running statistics, two-stream towers) with AVID_CUDA_GRAPH=1 and =0."""
import pytest
import torch

from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _train(monkeypatch, graphs, steps=6):
    from avid_cma_b200 import models, ops, optim
    from avid_cma_b200.criterions import AVID
    from avid_cma_b200.models._tower import _MATH
    monkeypatch.setenv("AVID_CUDA_GRAPH", "1" if graphs else "0")
    N, K, B = 64, 32, 2
    model = models.av_wrapper('R2Plus1D', {'depth': 18}, 'Conv2D', {'depth': 10}, proj_dim=[512, 512, 128])
    model.load_state_dict(synth.fill_state_dict(model.state_dict(), seed=13))
    model.video_model.math = model.audio_model.math = _MATH["bf16x3"]
    model = model.to(DEV).train()
    torch.manual_seed(5)
    crit = AVID(num_data=N, embedding_dim=128, num_negatives=K, momentum=0.5, xModal_coeff=1., wModal_coeff=0., device=0)
    crit.nce_average.view1_mem.copy_(synth.bank(N, seed=13, tag="bank_v"))
    crit.nce_average.view2_mem.copy_(synth.bank(N, seed=13, tag="bank_a"))
    opt = optim.Adam(model.parameters(), lr=2e-4, weight_decay=1e-5)
    losses = []
    ops.reset_launch_count()
    per_step = []
    for i in range(steps):
        video, audio = synth.clips(B, 4, 32, seed=100 + i).to(DEV), synth.spectrograms(B, 40, 33, seed=100 + i).to(DEV)
        y = synth.instance_ids(B, N, seed=100 + i).to(DEV)
        idx = synth.negatives(y.cpu(), K, N, seed=100 + i).to(DEV)
        crit.nce_average.sample_negatives = lambda y_, K_, idx=idx: idx
        n0 = ops.launch_count()
        ve, ae = model(video, audio)
        loss, _ = crit(ve, ae, y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss))
        per_step.append(ops.launch_count() - n0)
    torch.cuda.synchronize()
    graphed = [t.__dict__.get('_graphs') for t in (model.video_model, model.audio_model)]
    return losses, {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}, per_step, graphed


def test_graphed_towers_equal_eager_towers(monkeypatch):
    l0, sd0, n0, g0 = _train(monkeypatch, graphs=False)
    l1, sd1, n1, g1 = _train(monkeypatch, graphs=True)
    assert all(g is None or all(e[1] is None for e in g.values()) for g in g0)                       # eager run: nothing captured
    assert all(g and all(e[1] not in (None, False) and e[1].bwd is not None for e in g.values()) for g in g1), "towers were not captured"
    assert n1[-1] == n0[-1] and n1[0] == n0[0], (n0, n1)                 # replayed launches are counted like direct ones
    for a, b in zip(l0, l1):
        assert abs(a - b) <= 2e-5 * abs(a), (l0, l1)                   # fp32 atomics reorder sums: not bit-identical
    worst = 0.0
    for k in sd0:
        a, b = sd0[k].double(), sd1[k].double()
        if a.numel() and a.is_floating_point():
            worst = max(worst, float((a - b).norm() / a.norm().clamp_min(1e-30)))
        else:
            assert torch.equal(sd0[k], sd1[k]), k                      # num_batches_tracked
    assert worst < 2e-3, worst      # Adam's +-lr steps amplify last-bit gradient differences of near-zero gradients (see smoke())
