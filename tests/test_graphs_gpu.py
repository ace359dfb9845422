"""CUDA-graph replay of the towers (models/_tower.py::GraphedTower) == eager launches, on the SAME parameters and inputs:
calls 1-2 of a tower run eagerly, call 3 captures, later calls replay.  (Two separate training runs cannot be compared step by
step: fp32 atomics reorder the sums, and train-mode BatchNorm at batch 2 plus Adam's +-lr first step amplify that to per cent.)"""
import pytest
import torch

from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_graphed_towers_equal_eager_towers(monkeypatch):
    from avid_cma_b200 import models, ops
    from avid_cma_b200.models._tower import _MATH
    monkeypatch.setenv("AVID_CUDA_GRAPH", "1")
    B = 2
    model = models.av_wrapper('R2Plus1D', {'depth': 18}, 'Conv2D', {'depth': 10}, proj_dim=[512, 512, 128])
    model.load_state_dict(synth.fill_state_dict(model.state_dict(), seed=13))
    model.video_model.math = model.audio_model.math = _MATH["bf16x3"]
    model = model.to(DEV).train()
    wv, wa = synth.normal((B, 128), 1, "wv").to(DEV), synth.normal((B, 128), 1, "wa").to(DEV)
    keys = ["video_model.conv1.0.weight", "video_model.conv2x.0.spt_conv1.weight", "video_model.conv5x.1.out_bn.weight",
            "audio_model.block2.conv1.weight", "audio_model.conv1.1.bias", "video_proj.projection.0.weight"]
    params = dict(model.named_parameters())

    def run(seed, graphs):
        monkeypatch.setenv("AVID_CUDA_GRAPH", "1" if graphs else "0")
        video, audio = synth.clips(B, 4, 32, seed=seed).to(DEV), synth.spectrograms(B, 40, 33, seed=seed).to(DEV)
        for p in model.parameters():
            p.grad = None
        n0 = ops.launch_count()
        ve, ae = model(video, audio)
        loss = (ve * wv).sum() + (ae * wa).sum()
        loss.backward()
        torch.cuda.synchronize()
        return float(loss), ve.detach().clone(), ae.detach().clone(), {k: params[k].grad.detach().clone() for k in keys}, ops.launch_count() - n0

    eager = {s: run(s, graphs=False) for s in (100, 101)}          # eager references (these calls do not count towards the warm-up)
    first = [run(100, graphs=True) for _ in range(2)]              # calls 1-2 with graphs enabled: still eager
    captured = run(100, graphs=True)                               # call 3: capture + first replay
    replay_same = run(100, graphs=True)                            # replay, same input
    replay_other = run(101, graphs=True)                           # replay, fresh input copied into the static buffer
    graphs = [t.__dict__.get('_graphs') for t in (model.video_model, model.audio_model)]
    assert all(g and any(e[1] not in (None, False) and e[1].bwd is not None for e in g.values()) for g in graphs), "towers were not captured"
    for got, want in ((first[0], eager[100]), (captured, eager[100]), (replay_same, eager[100]), (replay_other, eager[101])):
        assert abs(got[0] - want[0]) <= 1e-4 * abs(want[0]), (got[0], want[0])
        assert _rel(got[1], want[1]) < 1e-4 and _rel(got[2], want[2]) < 1e-4
        for k in keys:
            assert _rel(got[3][k], want[3][k]) < 2e-3, (k, _rel(got[3][k], want[3][k]))     # atomics + ReLU-gate flips in early layers
        assert got[4] == want[4], (got[4], want[4])                # replayed launches are counted like direct ones
    # the replays really differ between inputs (the static input buffer is refreshed)
    assert abs(replay_other[0] - replay_same[0]) > 1e-3 * abs(replay_same[0])
    # BatchNorm bookkeeping advanced once per call, graphed or not
    assert int(model.video_model.conv1[1].num_batches_tracked) == 7
