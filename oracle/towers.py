"""Oracle for the encoder half of the hot path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Functional torch-CPU restatement, driven by a flat `state_dict` with the reference's
parameter names, of

  * R2Plus1D(depth=18).forward           models/video.py:44-54, network_blocks.py:30-60
  * Conv2D(depth=10).forward             models/audio.py:34-44, network_blocks.py:13-27
  * Head / AV_Wrapper.forward            models/av_wrapper.py:17-61

Train-mode BatchNorm uses batch statistics (biased variance, eps 1e-5) and updates the
running statistics in the passed dict with momentum 0.1 and the unbiased variance, exactly
as nn.BatchNorm{2,3}d does.
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1

VIDEO_STAGES = (("conv2x", 64, 64, 1), ("conv3x", 64, 128, 2), ("conv4x", 128, 256, 2), ("conv5x", 256, 512, 2))
AUDIO_BLOCKS = (("block1", 64, 64, 2), ("block2", 64, 128, 2), ("block3", 128, 256, 2), ("block4", 256, 512, 1))


def _bn(x, sd, prefix, training):
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training and prefix + ".num_batches_tracked" in sd:
        sd[prefix + ".num_batches_tracked"] += 1
    if training:
        # batch_norm wants running stats in the activation dtype; update copies and write back
        rm_c, rv_c = rm.to(x.dtype), rv.to(x.dtype)
        y = F.batch_norm(x, rm_c, rv_c, sd[prefix + ".weight"].to(x.dtype), sd[prefix + ".bias"].to(x.dtype),
                         True, BN_MOMENTUM, BN_EPS)
        if rm_c is not rm:
            rm.copy_(rm_c.detach())
            rv.copy_(rv_c.detach())
        return y
    return F.batch_norm(x, rm.to(x.dtype), rv.to(x.dtype), sd[prefix + ".weight"].to(x.dtype),
                        sd[prefix + ".bias"].to(x.dtype), False, BN_MOMENTUM, BN_EPS)


def _w(sd, key, x):
    return sd[key].to(x.dtype)


def r2p1d_block(x, sd, p, stride, has_res, training):
    """BasicR2P1DBlock.forward (network_blocks.py:53-60)."""
    s = stride
    h = F.conv3d(x, _w(sd, p + ".spt_conv1.weight", x), stride=(1, s, s), padding=(0, 1, 1))
    h = F.relu(_bn(h, sd, p + ".spt_bn1", training))
    h = F.conv3d(h, _w(sd, p + ".tmp_conv1.weight", x), stride=(s, 1, 1), padding=(1, 0, 0))
    h = F.relu(_bn(h, sd, p + ".tmp_bn1", training))
    h = F.conv3d(h, _w(sd, p + ".spt_conv2.weight", x), padding=(0, 1, 1))
    h = F.relu(_bn(h, sd, p + ".spt_bn2", training))
    h = F.conv3d(h, _w(sd, p + ".tmp_conv2.weight", x), padding=(1, 0, 0))
    r = F.conv3d(x, _w(sd, p + ".res_conv.weight", x), stride=(s, s, s)) if has_res else x
    return F.relu(_bn(h + r, sd, p + ".out_bn", training))


def video_tower(x, sd, prefix="video_model", training=True, return_embs=False):
    """R2Plus1D depth 18 (video.py:18-54). x: (B,3,T,H,W)."""
    p = prefix
    h = F.conv3d(x, _w(sd, p + ".conv1.0.weight", x), stride=(1, 2, 2), padding=(1, 3, 3))
    h = F.relu(_bn(h, sd, p + ".conv1.1", training))
    h = F.max_pool3d(h, (1, 3, 3), (1, 2, 2), (0, 1, 1))
    embs = {"conv1": h}
    for name, cin, cout, stride in VIDEO_STAGES:
        h = r2p1d_block(h, sd, f"{p}.{name}.0", stride, cin != cout or stride != 1, training)
        h = r2p1d_block(h, sd, f"{p}.{name}.1", 1, False, training)
        embs[name] = h
    pooled = F.adaptive_max_pool3d(h, 1)
    embs["pool"] = pooled
    return embs if return_embs else pooled


def audio_tower(x, sd, prefix="audio_model", training=True, return_embs=False):
    """Conv2D depth 10 (audio.py:20-44). x: (B,1,T,F)."""
    p = prefix
    h = F.conv2d(x, _w(sd, p + ".conv1.0.weight", x), stride=2, padding=3)
    h = F.relu(_bn(h, sd, p + ".conv1.1", training))
    embs = {}
    for (name, cin, cout, stride), tag in zip(AUDIO_BLOCKS, ("conv2x", "conv3x", "conv4x", "conv5x")):
        q = f"{p}.{name}"
        h = F.conv2d(h, _w(sd, q + ".conv1.weight", x), stride=stride, padding=1)
        h = F.relu(_bn(h, sd, q + ".bn1", training))
        h = F.conv2d(h, _w(sd, q + ".conv2.weight", x), padding=1)
        h = F.relu(_bn(h, sd, q + ".bn2", training))
        embs[tag] = h
    pooled = F.adaptive_max_pool2d(h, 1)
    embs["pool"] = pooled
    return embs if return_embs else pooled


def head(x, sd, prefix):
    """Head.forward (av_wrapper.py:17-33): Linear(+ReLU) chain `prefix.projection.{0,2,4,...}`."""
    idx = sorted({int(k[len(prefix) + len(".projection."):].split(".")[0]) for k in sd if k.startswith(prefix + ".projection.")})
    for n, i in enumerate(idx):
        x = F.linear(x, _w(sd, f"{prefix}.projection.{i}.weight", x), _w(sd, f"{prefix}.projection.{i}.bias", x))
        if n < len(idx) - 1:
            x = F.relu(x)
    return x


def av_forward(video, audio, sd, training=True):
    """AV_Wrapper.forward (av_wrapper.py:50-61) -> (video_emb, audio_emb)."""
    v = video_tower(video, sd, training=training)
    v = head(v.view(v.shape[0], v.shape[1]), sd, "video_proj")
    a = audio_tower(audio, sd, training=training)
    a = head(a.view(a.shape[0], a.shape[1]), sd, "audio_proj")
    return v, a


def param_keys(sd):
    """Trainable entries of a reference state_dict (everything but BN running statistics)."""
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"))]


def state_dict_template(proj_dim=(512, 512, 128)):
    """Names and shapes of the reference `av_wrapper('R2Plus1D',{depth:18},'Conv2D',{depth:10},proj_dim)`
    state_dict (267 entries), built without importing the reference."""
    from collections import OrderedDict
    sd = OrderedDict()

    def bn(p, c):
        sd[p + ".weight"] = torch.ones(c)
        sd[p + ".bias"] = torch.zeros(c)
        sd[p + ".running_mean"] = torch.zeros(c)
        sd[p + ".running_var"] = torch.ones(c)
        sd[p + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    sd["video_model.conv1.0.weight"] = torch.zeros(64, 3, 3, 7, 7)
    bn("video_model.conv1.1", 64)
    for name, cin, cout, stride in VIDEO_STAGES:
        for b in (0, 1):
            p = f"video_model.{name}.{b}"
            ci = cin if b == 0 else cout
            sd[p + ".spt_conv1.weight"] = torch.zeros(cout, ci, 1, 3, 3)
            bn(p + ".spt_bn1", cout)
            sd[p + ".tmp_conv1.weight"] = torch.zeros(cout, cout, 3, 1, 1)
            bn(p + ".tmp_bn1", cout)
            sd[p + ".spt_conv2.weight"] = torch.zeros(cout, cout, 1, 3, 3)
            bn(p + ".spt_bn2", cout)
            sd[p + ".tmp_conv2.weight"] = torch.zeros(cout, cout, 3, 1, 1)
            bn(p + ".out_bn", cout)
            if b == 0 and (cin != cout or stride != 1):
                sd[p + ".res_conv.weight"] = torch.zeros(cout, cin, 1, 1, 1)
    sd["audio_model.conv1.0.weight"] = torch.zeros(64, 1, 7, 7)
    bn("audio_model.conv1.1", 64)
    for name, cin, cout, stride in AUDIO_BLOCKS:
        p = f"audio_model.{name}"
        sd[p + ".conv1.weight"] = torch.zeros(cout, cin, 3, 3)
        bn(p + ".bn1", cout)
        sd[p + ".conv2.weight"] = torch.zeros(cout, cout, 3, 3)
        bn(p + ".bn2", cout)
    for tower in ("video_proj", "audio_proj"):
        d = 512
        for i, o in enumerate(proj_dim):
            sd[f"{tower}.projection.{2 * i}.weight"] = torch.zeros(o, d)
            sd[f"{tower}.projection.{2 * i}.bias"] = torch.zeros(o)
            d = o
    return sd
