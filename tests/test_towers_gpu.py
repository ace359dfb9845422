"""GPU parity of the encoder kernels (conv fwd/dgrad/wgrad, BN, pools, heads, Adam) and of the whole
training step at BASELINE config 1, through the C ABI, against torch fp64 on CPU / the golden vectors."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import criterion as oc
from oracle import synth, towers

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _close(a, b, rtol, atol=0.0):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def _rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


# (n, ci, co, (t,h,w), kernel, stride, padding): every layer type of both towers, small extents, ragged pixel counts
CONV_CASES = [
    (2, 3, 64, (4, 18, 22), (3, 7, 7), (1, 2, 2), (1, 3, 3)),      # video stem (ci padded 3 -> 4)
    (3, 1, 64, (1, 21, 27), (1, 7, 7), (1, 2, 2), (0, 3, 3)),      # audio stem (ci padded 1 -> 4)
    (2, 64, 64, (3, 9, 11), (1, 3, 3), (1, 1, 1), (0, 1, 1)),      # spatial
    (2, 64, 64, (4, 5, 7), (3, 1, 1), (1, 1, 1), (1, 0, 0)),       # temporal
    (2, 64, 128, (4, 10, 10), (1, 3, 3), (1, 2, 2), (0, 1, 1)),    # strided spatial (stage entry)
    (2, 128, 128, (4, 5, 5), (3, 1, 1), (2, 1, 1), (1, 0, 0)),     # strided temporal
    (2, 64, 128, (4, 10, 10), (1, 1, 1), (2, 2, 2), (0, 0, 0)),    # residual 1x1x1 s2
    (1, 256, 512, (1, 7, 9), (1, 3, 3), (1, 1, 1), (0, 1, 1)),     # audio block4-like
    (5, 128, 256, (1, 13, 17), (1, 3, 3), (1, 2, 2), (0, 1, 1)),   # audio block3 entry
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[f"c{i}" for i in range(len(CONV_CASES))])
def test_conv_forward_dgrad_wgrad(case):
    from avid_cma_b200 import ops
    from avid_cma_b200.models.network_blocks import pad_channels
    n, ci, co, (t, h, w), k, s, p = case
    g = torch.Generator().manual_seed(hash(case) % 2 ** 31)
    x = torch.randn(n, ci, t, h, w, generator=g)
    wt = torch.randn(co, ci, *k, generator=g) / (ci * k[0] * k[1] * k[2]) ** 0.5
    xd, wd = x.double().requires_grad_(True), wt.double().requires_grad_(True)
    ref = F.conv3d(xd, wd, stride=s, padding=p)
    dout = torch.randn(ref.shape, generator=g)
    addend = torch.randn(ref.shape, generator=g)
    ref.backward(dout.double())
    cp = pad_channels(ci)
    xc = ops.nchw_to_nhwc(x.to(DEV), c_pad=cp)
    assert xc.shape == (n, t, h, w, cp)
    w_tap, w_tap_t = ops.filter_to_tapmajor(wt.to(DEV), ci_pad=cp)
    shape = ops.conv_shape(n, t, h, w, cp, co, k, s, p)
    out = ops.conv_forward(shape, xc, w_tap)
    assert _rel(ops.nhwc_to_nchw(out), ref) < 2e-6
    add_c = ops.nchw_to_nhwc(addend.to(DEV))
    out2 = ops.conv_forward(shape, xc, w_tap, addend=add_c)
    assert _rel(ops.nhwc_to_nchw(out2), ref + addend.double()) < 2e-6
    dout_c = ops.nchw_to_nhwc(dout.to(DEV))
    dw = ops.filter_from_tapmajor(ops.conv_wgrad(shape, xc, dout_c), wt.to(DEV))
    assert _rel(dw, wd.grad) < 5e-6
    if ci >= 64:
        din = ops.conv_dgrad(shape, dout_c, w_tap_t)
        assert _rel(ops.nhwc_to_nchw(din), xd.grad) < 2e-6
        xadd = torch.randn(x.shape, generator=g)
        din2 = ops.conv_dgrad(shape, dout_c, w_tap_t, addend=ops.nchw_to_nhwc(xadd.to(DEV)))
        assert _rel(ops.nhwc_to_nchw(din2), xd.grad + xadd.double()) < 2e-6


def test_layout_round_trips():
    from avid_cma_b200 import ops
    x = torch.randn(3, 64, 2, 5, 7, device=DEV)
    assert torch.equal(ops.nhwc_to_nchw(ops.nchw_to_nhwc(x)), x)
    assert torch.equal(ops.nchw_to_nhwc(x), x.permute(0, 2, 3, 4, 1).contiguous())
    w = torch.randn(128, 64, 3, 1, 1, device=DEV)
    w_tap, w_tap_t = ops.filter_to_tapmajor(w)
    assert torch.equal(w_tap, w.reshape(128, 64, 3).permute(2, 1, 0).contiguous())
    assert torch.equal(w_tap_t, w.reshape(128, 64, 3).permute(2, 0, 1).contiguous())
    assert torch.equal(ops.filter_from_tapmajor(w_tap, w), w)


@pytest.mark.parametrize("c,rows", [(64, 4 * 8 * 28 * 28), (512, 37), (128, 1000)])
def test_batchnorm_relu_forward_backward(c, rows):
    from avid_cma_b200 import ops
    g = torch.Generator().manual_seed(c + rows)
    x = torch.randn(rows, c, generator=g) * 2 + 0.5
    gamma, beta = 1 + 0.2 * torch.randn(c, generator=g), 0.3 * torch.randn(c, generator=g)
    rm, rv = torch.randn(c, generator=g), torch.rand(c, generator=g) + 0.5
    dy = torch.randn(rows, c, generator=g)
    xd = x.double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    rm_d, rv_d = rm.double().clone(), rv.double().clone()
    # channels-first view for F.batch_norm: (rows, c) is already (N, C)
    ref = F.relu(F.batch_norm(xd, rm_d, rv_d, gd, bd, True, 0.1, 1e-5))
    ref.backward(dy.double())
    xg, rmg, rvg = x.to(DEV), rm.to(DEV), rv.to(DEV)
    st = ops.bn_train_stats(xg, gamma.to(DEV), beta.to(DEV), rmg, rvg)
    y = ops.bn_relu_forward(xg, st.scale, st.shift)
    _close(y, ref, 1e-5, 1e-5)
    _close(rmg, rm_d, 1e-5, 1e-6)
    _close(rvg, rv_d, 1e-5, 1e-6)
    dx, dgamma, dbeta = ops.bn_relu_backward(xg, dy.to(DEV), st, gamma.to(DEV), beta.to(DEV))
    assert _rel(dx, xd.grad) < 2e-5
    assert _rel(dgamma, gd.grad) < 1e-5 and _rel(dbeta, bd.grad) < 1e-5


def test_pools_forward_backward():
    from avid_cma_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 64, 3, 13, 10, generator=g)
    xd = x.double().requires_grad_(True)
    ref = F.max_pool3d(xd, (1, 3, 3), (1, 2, 2), (0, 1, 1))
    dy = torch.randn(ref.shape, generator=g)
    ref.backward(dy.double())
    xc = ops.nchw_to_nhwc(x.to(DEV))
    y, amax = ops.maxpool_1x3x3_forward(xc)
    _close(ops.nhwc_to_nchw(y), ref, 0, 0)
    dx = ops.maxpool_1x3x3_backward(amax, ops.nchw_to_nhwc(dy.to(DEV)), xc.shape)
    _close(ops.nhwc_to_nchw(dx), xd.grad, 1e-6, 1e-6)
    # ties: after ReLU half the inputs are exact zeros; the gradient must go to the FIRST maximum of a window, like ATen
    xr = torch.relu(x).double().requires_grad_(True)
    refr = F.max_pool3d(xr, (1, 3, 3), (1, 2, 2), (0, 1, 1))
    refr.backward(dy.double())
    yr, amr = ops.maxpool_1x3x3_forward(ops.nchw_to_nhwc(torch.relu(x).to(DEV)))
    dxr = ops.maxpool_1x3x3_backward(amr, ops.nchw_to_nhwc(dy.to(DEV)), xc.shape)
    _close(ops.nhwc_to_nchw(yr), refr, 0, 0)
    _close(ops.nhwc_to_nchw(dxr), xr.grad, 1e-6, 1e-6)
    xd2 = x.double().requires_grad_(True)
    ref2 = F.adaptive_max_pool3d(xd2, 1).flatten(1)
    dy2 = torch.randn(ref2.shape, generator=g)
    ref2.backward(dy2.double())
    y2, am = ops.global_maxpool_forward(xc)
    _close(y2, ref2, 0, 0)
    dx2 = ops.global_maxpool_backward(dy2.to(DEV), am, xc.shape)
    _close(ops.nhwc_to_nchw(dx2), xd2.grad, 0, 0)


def test_head_forward_backward():
    from avid_cma_b200.models.av_wrapper import Head
    torch.manual_seed(3)
    head = Head(512, [512, 512, 128]).to(DEV)
    x = torch.randn(7, 512, device=DEV, requires_grad=True)
    y = head(x)
    dy = torch.randn_like(y)
    y.backward(dy)
    ref_head = torch.nn.Sequential(*[m for m in head.projection]).double().cpu()
    xr = x.detach().cpu().double().requires_grad_(True)
    yr = xr
    for m in ref_head:
        yr = m(yr) if isinstance(m, torch.nn.Linear) else F.relu(yr)
    yr.backward(dy.cpu().double())
    assert _rel(y, yr) < 1e-5 and _rel(x.grad, xr.grad) < 1e-5
    for (n, p), (_, pr) in zip(head.named_parameters(), ref_head.named_parameters()):
        assert _rel(p.grad, pr.grad) < 1e-5, n


def test_adam_step_matches_torch():
    from avid_cma_b200 import ops
    torch.manual_seed(0)
    p = torch.randn(10007, device=DEV)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=2e-4, weight_decay=1e-5)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        g = torch.randn_like(p)
        ref.grad = g.clone()
        opt.step()
        ops.adam_step_(p, g, m, v, step, 2e-4, (0.9, 0.999), 1e-8, 1e-5)
        _close(p, ref, 1e-5, 1e-7)


def test_adam_multi_tensor_matches_torch():
    """avid_adam_step_multi over 70 tensors of ragged sizes (3 launches of <= 32 tensors) == torch.optim.Adam."""
    from avid_cma_b200 import optim
    g = torch.Generator().manual_seed(3)
    sizes = [1, 7, 4096, 4097, 64 * 64 * 9, 512] * 11 + [100003, 3, 12289, 5]
    ours = [torch.randn(n, generator=g).to(DEV).requires_grad_(True) for n in sizes]
    ref = [p.detach().clone().requires_grad_(True) for p in ours]
    o1, o2 = optim.Adam(ours, lr=2e-4, weight_decay=1e-5), torch.optim.Adam(ref, lr=2e-4, weight_decay=1e-5)
    for _ in range(3):
        for a, b in zip(ours, ref):
            a.grad = torch.randn(a.shape, generator=g).to(DEV)
            b.grad = a.grad.clone()
        o1.step()
        o2.step()
    for a, b in zip(ours, ref):
        _close(a, b, 1e-5, 1e-7)


def test_filter_to_planes_matches_tapmajor_split():
    from avid_cma_b200 import ops
    w = torch.randn(128, 64, 1, 3, 3, generator=torch.Generator().manual_seed(2)).to(DEV)
    (f_hi, f_lo), (d_hi, d_lo) = ops.filter_to_planes(w)
    w_tap, w_tap_t = ops.filter_to_tapmajor(w)
    for got, want in ((f_hi, f_lo), ops.split_bf16(w_tap_t)), ((d_hi, d_lo), ops.split_bf16(w_tap)):
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])


def _load_model(seed=0, math="fp32"):
    from avid_cma_b200 import models
    from avid_cma_b200.models._tower import _MATH
    model = models.av_wrapper('R2Plus1D', {'depth': 18}, 'Conv2D', {'depth': 10}, proj_dim=[512, 512, 128])
    model.load_state_dict(synth.fill_state_dict(model.state_dict(), seed=seed))
    model.video_model.math = model.audio_model.math = _MATH[math]
    return model.to(DEV).train()


@pytest.mark.parametrize("math", ["fp32", "bf16x3"])
def test_training_step_config1_matches_reference_golden(golden, math):
    """BASELINE config 1: batch 4, 8x3x112x112 + 1x100x129, bank 64, K = 1024, injected negatives: embeddings, loss,
    every parameter-gradient norm, selected gradients, BN running stats and updated bank rows vs the imported reference."""
    from avid_cma_b200.criterions import AVID
    g = golden("step_config1")
    B, N, K, size, seed = int(g["B"]), int(g["N"]), int(g["K"]), int(g["size"]), int(g["seed"])
    spec = g["spec"].tolist()
    model = _load_model(seed, math)
    crit = AVID(num_data=N, embedding_dim=128, num_negatives=K, momentum=0.5, xModal_coeff=1., wModal_coeff=0., device=0)
    crit.nce_average.view1_mem.copy_(synth.bank(N, seed=seed, tag="bank_v"))
    crit.nce_average.view2_mem.copy_(synth.bank(N, seed=seed, tag="bank_a"))
    y = torch.from_numpy(g["y"])
    idx = synth.negatives(y, K, N, seed).to(DEV)
    crit.nce_average.sample_negatives = lambda y_, K_: idx
    video, audio = synth.clips(B, 8, size, seed).to(DEV), synth.spectrograms(B, spec[0], spec[1], seed).to(DEV)
    ve, ae = model(video, audio)
    loss, log = crit(ve, ae, y.to(DEV))
    loss.backward()
    tol = 1e-3   # north-star tolerance: 1e-3 relative on fp32 embeddings and loss
    assert _rel(ve, torch.from_numpy(g["video_emb"])) < tol
    assert _rel(ae, torch.from_numpy(g["audio_emb"])) < tol
    _close(loss, g["total"], tol)
    _close(crit.criterion.avg_exp_score, g["Z"], tol)
    _close(log["Loss/v2a"], g["Loss/v2a"], tol)
    norms = dict(zip(g["grad_names"].tolist(), g["grad_norms"].tolist()))
    worst = 0.0
    for n, p in model.named_parameters():
        assert p.grad is not None, n
        worst = max(worst, abs(float(p.grad.double().norm()) - norms[n]) / norms[n])
    assert worst < (5e-3 if math == "fp32" else 2e-2), worst      # bf16x3: ReLU-gate flips, see the per-tensor check below
    params = dict(model.named_parameters())
    sd = model.state_dict()
    for k in g.files:
        if k.startswith("grad::"):
            # early-layer gradients pass through 30+ train-mode BNs at batch 4: the fp32 reference itself is ~4e-3 away
            # from the fp64 result there, so the bar is "no further from fp64 than 2x the reference's own rounding"
            truth = torch.from_numpy(g["fp64::" + k])
            ref_err = _rel(torch.from_numpy(g[k]), truth)
            # bf16x3 carries ~16 significand bits per operand: embeddings move by ~1e-5, which flips a handful of ReLU gates
            # (4 of ~1M in the audio tower, scripts/emulate_bf16x3_budget.py); every flip is an O(1) change of one gradient
            # element, i.e. ~sqrt(flip fraction) = several 1e-3 in relative L2 on the layers below it -- the same mechanism
            # that puts the fp32 reference 4e-3 from fp64 on the (much larger) video tower.
            floor = 1e-3 if math == "fp32" else 2e-2
            assert _rel(params[k[6:]].grad, truth) < max(floor, (2 if math == "fp32" else 8) * ref_err), (k, ref_err)
        elif k.startswith("grad_slice::"):
            want = torch.from_numpy(g[k])
            # fp32 reference vs fp64 is already 4e-3 here (ReLU-gate flips, see above); 16 operand bits flip ~256x more gates
            assert _rel(params[k[12:]].grad[:want.shape[0]], want) < (1e-2 if math == "fp32" else 5e-2), k
        elif k.startswith("rm::"):
            _close(sd[k[4:] + ".running_mean"], g[k], 1e-3, 1e-6)
        elif k.startswith("rv::"):
            _close(sd[k[4:] + ".running_var"], g[k], 1e-3, 1e-6)
    assert int(sd["video_model.conv1.1.num_batches_tracked"]) == 1
    assert _rel(crit.nce_average.view1_mem[y.to(DEV)], torch.from_numpy(g["rows_v"])) < tol
    assert _rel(crit.nce_average.view2_mem[y.to(DEV)], torch.from_numpy(g["rows_a"])) < tol


def test_training_step_config1_bf16_single_pass(golden):
    """The documented fast mode (one bf16 MMA per product): exercised end to end; with train-mode BN at batch 4 it is
    NOT expected to meet the 1e-3 parity bar (SURVEY.md §7 measured 5.5e-2 for bf16 autocast on the reference itself)."""
    from avid_cma_b200.criterions import AVID
    g = golden("step_config1")
    B, N, K, size, seed = int(g["B"]), int(g["N"]), int(g["K"]), int(g["size"]), int(g["seed"])
    spec = g["spec"].tolist()
    model = _load_model(seed, "bf16")
    crit = AVID(num_data=N, embedding_dim=128, num_negatives=K, momentum=0.5, xModal_coeff=1., wModal_coeff=0., device=0)
    crit.nce_average.view1_mem.copy_(synth.bank(N, seed=seed, tag="bank_v"))
    crit.nce_average.view2_mem.copy_(synth.bank(N, seed=seed, tag="bank_a"))
    y = torch.from_numpy(g["y"])
    idx = synth.negatives(y, K, N, seed).to(DEV)
    crit.nce_average.sample_negatives = lambda y_, K_: idx
    ve, ae = model(synth.clips(B, 8, size, seed).to(DEV), synth.spectrograms(B, spec[0], spec[1], seed).to(DEV))
    loss, _ = crit(ve, ae, y.to(DEV))
    loss.backward()
    ev, ea = _rel(ve, torch.from_numpy(g["video_emb"])), _rel(ae, torch.from_numpy(g["audio_emb"]))
    print("bf16 single-pass embedding errors", ev, ea)
    assert ev < 0.25 and ea < 0.1
    assert abs(float(loss) - float(g["total"])) < 0.05 * float(g["total"])
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())


def test_towers_eval_mode_and_return_embs():
    """Inference path: BN uses running statistics; return_embs gives the reference's NC(D)HW taps (video.py:51-52)."""
    model = _load_model(1).eval()
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    video, audio = synth.clips(2, 4, 32, 9), synth.spectrograms(2, 40, 33, 9)
    with torch.no_grad():
        taps_v = model.video_model(video.to(DEV), return_embs=True)
        taps_a = model.audio_model(audio.to(DEV), return_embs=True)
        ve, ae = model(video.to(DEV), audio.to(DEV))
    ref_v = towers.video_tower(video.double(), sd, training=False, return_embs=True)
    ref_a = towers.audio_tower(audio.double(), sd, training=False, return_embs=True)
    assert set(taps_v) == set(ref_v) and set(taps_a) == set(ref_a)
    for k in ref_v:
        assert taps_v[k].shape == ref_v[k].shape and _rel(taps_v[k], ref_v[k]) < 1e-4, k
    for k in ref_a:
        assert taps_a[k].shape == ref_a[k].shape and _rel(taps_a[k], ref_a[k]) < 1e-4, k
    rv, ra = towers.av_forward(video.double(), audio.double(), sd, training=False)
    assert _rel(ve, rv) < 1e-4 and _rel(ae, ra) < 1e-4


@pytest.mark.parametrize("math", ["fp32", "bf16x3"])
def test_eval_mode_backward_matches_frozen_batchnorm(math):
    """model.eval() with gradients enabled (frozen-BatchNorm fine-tuning): BatchNorm is a per-channel affine map, so the input
    gradient has no batch-statistic terms while dgamma / dbeta are still the reduced sums (nn.BatchNorm in eval mode)."""
    model = _load_model(2, math).eval()
    sd = {k: v.detach().cpu().double().clone() for k, v in model.state_dict().items()}
    # running statistics away from (0, 1) so that a train-mode backward would be visibly different
    g = torch.Generator().manual_seed(5)
    for k in sd:
        if k.endswith("running_mean"):
            sd[k] = 0.2 * torch.randn(sd[k].shape, generator=g, dtype=torch.float64)
        elif k.endswith("running_var"):
            sd[k] = 0.5 + torch.rand(sd[k].shape, generator=g, dtype=torch.float64)
    model.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in sd.items()})
    keys = towers.param_keys(sd)
    for k in keys:
        sd[k].requires_grad_(True)
    video, audio = synth.clips(2, 4, 32, 11), synth.spectrograms(2, 40, 33, 11)
    wv, wa = synth.normal((2, 128), 11, "wv").double(), synth.normal((2, 128), 11, "wa").double()
    rv, ra = towers.av_forward(video.double(), audio.double(), sd, training=False)
    ((rv * wv).sum() + (ra * wa).sum()).backward()
    ve, ae = model(video.to(DEV), audio.to(DEV))
    ((ve * wv.float().to(DEV)).sum() + (ae * wa.float().to(DEV)).sum()).backward()
    tol = 1e-4 if math == "fp32" else 2e-3
    assert _rel(ve, rv) < tol and _rel(ae, ra) < tol
    params = dict(model.named_parameters())
    worst = max(_rel(params[k].grad, sd[k].grad) for k in keys)
    assert worst < (1e-3 if math == "fp32" else 2e-2), worst


def test_batched_filter_layout_conversions_equal_single_launches():
    """avid_filter_to_planes_multi / avid_filter_from_tapmajor_multi (one launch per tower) == the per-filter entry points."""
    from avid_cma_b200 import ops
    g = torch.Generator().manual_seed(4)
    shapes = [(64, 64, (1, 3, 3)), (128, 64, (1, 1, 1)), (128, 128, (3, 1, 1)), (512, 256, (3, 3)), (64, 3, (3, 7, 7))]
    ws = [torch.randn(co, ci, *k, generator=g).to(DEV) for co, ci, k in shapes]
    for need_lo in (True, False):
        single = [ops.filter_to_planes(w, need_lo) for w in ws[:4]]
        ops.prepare_filter_planes(ws[:4], need_lo)
        for w, want in zip(ws[:4], single):
            got = ops.filter_to_planes(w, need_lo)
            for a, b in zip(got[0] + got[1], want[0] + want[1]):
                assert (a is None and b is None) or torch.equal(a, b)
        assert not ops._plane_cache
    dws = [torch.randn(w[0, 0].numel(), 4 if w.shape[1] == 3 else w.shape[1], w.shape[0], generator=g).to(DEV) for w in ws]
    want = [ops.filter_from_tapmajor(d, w) for d, w in zip(dws, ws)]
    ops.defer_filter_gradients(True)
    got = [ops.filter_from_tapmajor(d, w) for d, w in zip(dws, ws)]
    ops.flush_filter_gradients()
    torch.cuda.synchronize()
    for a, b in zip(got, want):
        assert torch.equal(a, b)
