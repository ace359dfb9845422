"""GPU parity of the tcgen05 convolutions (TMA im2col + UMMA) against torch fp64: bf16x3 (hi/lo split, ~fp32 accuracy)
and single-pass bf16."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# (n, ci, co, (t,h,w), kernel, stride, padding)
TC_CASES = [
    (2, 64, 64, (3, 9, 11), (1, 3, 3), (1, 1, 1), (0, 1, 1)),      # spatial, ragged pixel count, rows wrap inside a tile
    (2, 64, 64, (4, 5, 7), (3, 1, 1), (1, 1, 1), (1, 0, 0)),       # temporal
    (2, 64, 128, (4, 10, 10), (1, 3, 3), (1, 2, 2), (0, 1, 1)),    # strided spatial (stage entry)
    (2, 128, 128, (4, 5, 5), (3, 1, 1), (2, 1, 1), (1, 0, 0)),     # strided temporal
    (2, 64, 128, (4, 10, 10), (1, 1, 1), (2, 2, 2), (0, 0, 0)),    # residual 1x1x1 s2
    (1, 256, 512, (1, 7, 9), (1, 3, 3), (1, 1, 1), (0, 1, 1)),     # audio block4-like (2-D)
    (3, 128, 256, (2, 14, 14), (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    (1, 512, 512, (1, 4, 4), (3, 1, 1), (1, 1, 1), (1, 0, 0)),     # conv5x temporal at config-1 size (T = 1)
    (4, 64, 64, (8, 28, 28), (1, 3, 3), (1, 1, 1), (0, 1, 1)),     # config-1 conv2x spatial: 196 tiles
    (2, 64, 64, (1, 25, 33), (1, 3, 3), (1, 2, 2), (0, 1, 1)),     # audio block entry, odd extents (ragged parity classes)
    (2, 64, 128, (3, 7, 9), (3, 3, 3), (2, 2, 2), (1, 1, 1)),      # full 3-D strided filter: 8 parity classes
    (1, 64, 64, (5, 6, 7), (3, 1, 1), (2, 1, 1), (0, 0, 0)),       # unpadded strided temporal (negative tap offsets)
    (2, 64, 64, (5, 8, 16), (3, 1, 1), (1, 1, 1), (1, 0, 0)),      # temporal, h*w % 64 == 0: frame-fastest k-block order of the filter gradient
    (2, 128, 128, (3, 8, 8), (3, 1, 1), (1, 1, 1), (1, 0, 0)),     # same with two channel blocks
    (3, 64, 64, (1, 20, 129), (1, 3, 3), (1, 1, 1), (0, 1, 1)),    # CTA-pair kernel, wide 2-D frames (audio block1 width): two-slot ring, odd frame count
    (1, 64, 64, (5, 7, 6), (1, 3, 3), (1, 1, 1), (0, 1, 1)),       # CTA-pair kernel, tiny frames: one tile per frame, odd frame count
]


def _rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("x3", [True, False], ids=["bf16x3", "bf16"])
@pytest.mark.parametrize("case", TC_CASES, ids=[f"c{i}" for i in range(len(TC_CASES))])
def test_conv_tc_forward_and_dgrad(case, x3):
    from avid_cma_b200 import ops
    n, ci, co, (t, h, w), k, s, p = case
    g = torch.Generator().manual_seed(hash(case) % 2 ** 31)
    x = torch.randn(n, ci, t, h, w, generator=g)
    wt = torch.randn(co, ci, *k, generator=g) / (ci * k[0] * k[1] * k[2]) ** 0.5
    xd, wd = x.double().requires_grad_(True), wt.double().requires_grad_(True)
    ref = F.conv3d(xd, wd, stride=s, padding=p)
    dout = torch.randn(ref.shape, generator=g)
    addend = torch.randn(ref.shape, generator=g)
    ref.backward(dout.double())
    tol = 3e-5 if x3 else 1e-2
    xc = ops.nchw_to_nhwc(x.to(DEV))
    w_tap, w_tap_t = ops.filter_to_tapmajor(wt.to(DEV))
    shape = ops.conv_shape(n, t, h, w, ci, co, k, s, p)
    x_hi, x_lo = ops.split_bf16(xc, x3)
    wf_hi, wf_lo = ops.split_bf16(w_tap_t, x3)            # forward: [taps, co, ci]
    out = ops.conv_forward_tc(shape, x_hi, x_lo, wf_hi, wf_lo)
    torch.cuda.synchronize()
    assert _rel(ops.nhwc_to_nchw(out), ref) < tol
    add_c = ops.nchw_to_nhwc(addend.to(DEV))
    stats = torch.zeros(2, co, dtype=torch.float64, device=DEV)
    out2 = ops.conv_forward_tc(shape, x_hi, x_lo, wf_hi, wf_lo, addend=add_c, bn_stats=stats)
    assert _rel(ops.nhwc_to_nchw(out2), ref + addend.double()) < tol
    # fused BatchNorm statistics = per-channel sum / sum of squares of what was stored
    flat = out2.double().reshape(-1, co)
    assert _rel(stats[0], flat.sum(0)) < 1e-5 and _rel(stats[1], (flat * flat).sum(0)) < 1e-5
    d_hi, d_lo = ops.split_bf16(ops.nchw_to_nhwc(dout.to(DEV)), x3)
    dw = ops.filter_from_tapmajor(ops.conv_wgrad_tc(shape, x_hi, x_lo, d_hi, d_lo), wt.to(DEV))
    assert _rel(dw, wd.grad) < tol
    wd_hi, wd_lo = ops.split_bf16(w_tap, x3)              # dgrad: [taps, ci, co]
    din = torch.full((n, t, h, w, ci), float("nan"), device=DEV)      # every pixel must be written, also by strided gradients
    ops.conv_dgrad_tc(shape, d_hi, d_lo, wd_hi, wd_lo, out=din)
    assert _rel(ops.nhwc_to_nchw(din), xd.grad) < tol
    if k != (1, 1, 1):                                    # an addend needs every stride-parity class to meet a filter tap
        add_in = torch.randn(x.shape, generator=g)
        din2 = ops.conv_dgrad_tc(shape, d_hi, d_lo, wd_hi, wd_lo, addend=ops.nchw_to_nhwc(add_in.to(DEV)))
        assert _rel(ops.nhwc_to_nchw(din2), xd.grad + add_in.double()) < tol


def test_split_bf16_reconstructs_16_bits():
    from avid_cma_b200 import ops
    x = torch.randn(4096, device=DEV) * 3
    hi, lo = ops.split_bf16(x)
    assert float(((hi.float() + lo.float()) - x).abs().max() / x.abs().max()) < 2 ** -15


# (n, ci, (t,h,w), kernel, stride, padding): the two stems at reduced sizes (co = 64)
STEM_CASES = [
    (2, 3, (4, 20, 36), (3, 7, 7), (1, 2, 2), (1, 3, 3)),      # video stem, tiles ragged in h and w
    (2, 1, (1, 37, 45), (1, 7, 7), (1, 2, 2), (0, 3, 3)),      # audio stem (2-D), odd extents
    (1, 3, (2, 64, 64), (3, 7, 7), (1, 2, 2), (1, 3, 3)),      # full 16-wide tiles, more tiles than one wave would need rows
    (3, 3, (8, 112, 112), (3, 7, 7), (1, 2, 2), (1, 3, 3)),    # config-1 video stem: 1176 tiles -> persistent loop, TMEM double buffer
    (2, 1, (1, 100, 129), (1, 7, 7), (1, 2, 2), (0, 3, 3)),    # config-1 audio stem
]


@pytest.mark.parametrize("x3", [True, False], ids=["bf16x3", "bf16"])
@pytest.mark.parametrize("case", STEM_CASES, ids=[f"s{i}" for i in range(len(STEM_CASES))])
def test_stem_tc_forward_and_wgrad(case, x3):
    from avid_cma_b200 import ops
    n, ci, (t, h, w), k, s, p = case
    co = 64
    g = torch.Generator().manual_seed(hash(case) % 2 ** 31)
    x = torch.randn(n, ci, t, h, w, generator=g)
    wt = torch.randn(co, ci, *k, generator=g) / (ci * k[0] * k[1] * k[2]) ** 0.5
    xd, wd = x.double().requires_grad_(True), wt.double().requires_grad_(True)
    ref = F.conv3d(xd, wd, stride=s, padding=p)
    dout = torch.randn(ref.shape, generator=g)
    ref.backward(dout.double())
    tol = 3e-5 if x3 else 1e-2
    shape = ops.conv_shape(n, t, h, w, ci, co, k, s, p)
    x_hi, x_lo = ops.stem_pack(x.to(DEV), 2 * shape.wo + 8, p[2], x3)
    w_hi, w_lo = ops.stem_filter_pack(wt.to(DEV), x3)
    stats = torch.zeros(2, co, dtype=torch.float64, device=DEV)
    out = ops.stem_forward_tc(shape, x_hi, x_lo, w_hi, w_lo, bn_stats=stats)
    torch.cuda.synchronize()
    assert _rel(ops.nhwc_to_nchw(out), ref) < tol
    flat = out.double().reshape(-1, co)
    assert _rel(stats[0], flat.sum(0)) < 1e-5 and _rel(stats[1], (flat * flat).sum(0)) < 1e-5
    d_hi, d_lo = ops.split_bf16(ops.nchw_to_nhwc(dout.to(DEV)), x3)
    dw_tap = ops.stem_wgrad_tc(shape, x_hi, x_lo, d_hi, d_lo)
    dw = ops.filter_from_tapmajor(dw_tap, wt.to(DEV))
    assert _rel(dw, wd.grad) < tol


# BASELINE config-2 layer geometries at full resolution (batch reduced to keep the fp64 checks cheap); checked through
# size-independent properties: the adjoint identities <conv(x), d> = <x, dgrad(d)> = <w, wgrad(x, d)> and a spot check of
# random output pixels against a direct fp64 evaluation of the convolution sum.
FULL_CASES = [
    (8, 64, 64, (8, 56, 56), (1, 3, 3), (1, 1, 1), (0, 1, 1)),       # conv2x spatial
    (8, 64, 64, (8, 56, 56), (3, 1, 1), (1, 1, 1), (1, 0, 0)),       # conv2x temporal
    (8, 64, 64, (1, 100, 129), (1, 3, 3), (1, 1, 1), (0, 1, 1)),     # audio block1 (CTA-pair kernel, two-slot ring)
    (8, 64, 128, (8, 56, 56), (1, 3, 3), (1, 2, 2), (0, 1, 1)),      # conv3x entry, strided
    (8, 128, 128, (8, 28, 28), (3, 1, 1), (2, 1, 1), (1, 0, 0)),     # conv3x strided temporal
    (4, 3, 64, (8, 224, 224), (3, 7, 7), (1, 2, 2), (1, 3, 3)),      # video stem
    (8, 1, 64, (1, 200, 257), (1, 7, 7), (1, 2, 2), (0, 3, 3)),      # audio stem
]


def _spot_check(x, wt, out_nchw, k, s, p, gen, points=256):
    n, ci, t, h, w = x.shape
    co = wt.shape[0]
    _, _, to, ho, wo = out_nchw.shape
    xp = F.pad(x.double(), (p[2], p[2], p[1], p[1], p[0], p[0]))
    idx = torch.stack([torch.randint(0, m, (points,), generator=gen) for m in (n, to, ho, wo)], 1)
    worst = 0.0
    for ni, ti, hi, wi in idx.tolist():
        patch = xp[ni, :, ti * s[0]:ti * s[0] + k[0], hi * s[1]:hi * s[1] + k[1], wi * s[2]:wi * s[2] + k[2]]
        ref = (wt.double() * patch.unsqueeze(0)).sum((1, 2, 3, 4))
        got = out_nchw[ni, :, ti, hi, wi].double().cpu()
        worst = max(worst, float((got - ref).norm() / ref.norm().clamp_min(1e-30)))
    return worst


@pytest.mark.parametrize("case", FULL_CASES, ids=[f"f{i}" for i in range(len(FULL_CASES))])
def test_conv_tc_full_resolution_properties(case):
    from avid_cma_b200 import ops
    n, ci, co, (t, h, w), k, s, p = case
    g = torch.Generator().manual_seed(1000 + ci + co)
    x = torch.randn(n, ci, t, h, w, generator=g)
    wt = torch.randn(co, ci, *k, generator=g) / (ci * k[0] * k[1] * k[2]) ** 0.5
    shape = ops.conv_shape(n, t, h, w, ci, co, k, s, p)
    dout = torch.randn(n, co, shape.to, shape.ho, shape.wo, generator=g)
    dc = ops.nchw_to_nhwc(dout.to(DEV))
    d_hi, d_lo = ops.split_bf16(dc)
    stem = ci < 64
    if stem:
        x_hi, x_lo = ops.stem_pack(x.to(DEV), 2 * shape.wo + 8, p[2])
        w_hi, w_lo = ops.stem_filter_pack(wt.to(DEV))
        out = ops.stem_forward_tc(shape, x_hi, x_lo, w_hi, w_lo)
        dw = ops.filter_from_tapmajor(ops.stem_wgrad_tc(shape, x_hi, x_lo, d_hi, d_lo), wt.to(DEV))
    else:
        xc = ops.nchw_to_nhwc(x.to(DEV))
        w_tap, w_tap_t = ops.filter_to_tapmajor(wt.to(DEV))
        x_hi, x_lo = ops.split_bf16(xc)
        wf_hi, wf_lo = ops.split_bf16(w_tap_t)
        wd_hi, wd_lo = ops.split_bf16(w_tap)
        out = ops.conv_forward_tc(shape, x_hi, x_lo, wf_hi, wf_lo)
        dw = ops.filter_from_tapmajor(ops.conv_wgrad_tc(shape, x_hi, x_lo, d_hi, d_lo), wt.to(DEV))
        din = ops.conv_dgrad_tc(shape, d_hi, d_lo, wd_hi, wd_lo)
    out_nchw = ops.nhwc_to_nchw(out)
    assert _spot_check(x, wt, out_nchw, k, s, p, g) < 3e-5
    lhs = float((out.double() * dc.double()).sum())                         # <conv(x), d>
    scale = float(out.double().norm() * dc.double().norm())
    assert abs(lhs - float((dw.double() * wt.to(DEV).double()).sum())) < 2e-5 * scale    # = <w, wgrad(x, d)>
    if not stem:
        assert abs(lhs - float((ops.nhwc_to_nchw(din).double() * x.to(DEV).double()).sum())) < 2e-5 * scale   # = <x, dgrad(d)>


# (n, ci, co, (t,h,w) of the INPUT, kernel, stride, padding of the main convolution, stride of the 1x1x1 residual convolution)
SUB_ADDEND_CASES = [
    (2, 64, 128, (4, 10, 10), (1, 3, 3), (1, 2, 2), (0, 1, 1), (2, 2, 2)),    # stage entry of the video tower: spatial stride below a (2,2,2) residual
    (2, 64, 64, (3, 9, 11), (1, 3, 3), (1, 2, 2), (0, 1, 1), (2, 2, 2)),      # odd extents: ragged classes, ceil-sized addend grid
    (1, 128, 128, (5, 6, 6), (1, 3, 3), (1, 1, 1), (0, 1, 1), (2, 1, 1)),     # stride-1 main convolution (one class), temporal residual stride
    (2, 64, 128, (3, 7, 9), (3, 3, 3), (2, 2, 2), (1, 1, 1), (1, 2, 2)),      # 8 parity classes
]


@pytest.mark.parametrize("case", SUB_ADDEND_CASES, ids=[f"a{i}" for i in range(len(SUB_ADDEND_CASES))])
def test_conv_tc_dgrad_subsampled_addend_equals_scattered_dense_addend(case):
    """avid_conv_dgrad_tc_sub: the input gradient of a strided 1x1x1 residual convolution stays on its output grid; adding it as a
    subsampled addend must give bit for bit what the dense, zero-filled addend gives -- output AND fused BatchNorm-backward sums."""
    from avid_cma_b200 import ops
    n, ci, co, (t, h, w), k, s, p, rs = case
    g = torch.Generator().manual_seed(hash(case) % 2 ** 31)
    shape = ops.conv_shape(n, t, h, w, ci, co, k, s, p)
    dout = torch.randn(n, shape.to, shape.ho, shape.wo, co, generator=g).to(DEV)
    wt = (torch.randn(k[0] * k[1] * k[2], ci, co, generator=g) / (co * k[0] * k[1] * k[2]) ** 0.5).to(DEV)       # dgrad operand [taps, ci, co]
    d_hi, d_lo = ops.split_bf16(dout, True)
    w_hi, w_lo = ops.split_bf16(wt, True)
    ta, ha, wa = -(-t // rs[0]), -(-h // rs[1]), -(-w // rs[2])
    sub = torch.randn(n, ta, ha, wa, ci, generator=g).to(DEV)
    dense = torch.zeros(n, t, h, w, ci, device=DEV)
    dense[:, ::rs[0], ::rs[1], ::rs[2]] = sub
    z = torch.randn(n, t, h, w, ci, generator=g).to(DEV)
    st = ops.BNState(ci, DEV)
    st.mean.copy_(torch.randn(ci, generator=g) * 0.1)
    st.invstd.copy_(torch.rand(ci, generator=g) + 0.5)
    gamma, beta = (torch.rand(ci, generator=g) + 0.5).to(DEV), (torch.randn(ci, generator=g) * 0.1).to(DEV)
    fuse_ok = all(kk >= ss for kk, ss in zip(k, s))
    outs = []
    for addend, stride in ((dense, None), (sub, rs)):
        sums = torch.zeros(2, ci, dtype=torch.float64, device=DEV)
        out = ops.conv_dgrad_tc(shape, d_hi, d_lo, w_hi, w_lo, addend=addend, addend_stride=stride,
                                bn_fuse=(z, st, gamma, beta, sums) if fuse_ok else None)
        torch.cuda.synchronize()
        outs.append((out, sums))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.allclose(outs[0][1], outs[1][1], rtol=1e-12, atol=0)          # fp64 atomics: the order of the CTAs' contributions may differ
    plain = ops.conv_dgrad_tc(shape, d_hi, d_lo, w_hi, w_lo)
    assert torch.equal(outs[1][0] - plain, dense) or _rel(outs[1][0] - plain, dense) < 1e-6
