"""Two-tower wrapper + projection heads (reference: models/av_wrapper.py)."""
import torch
import torch.nn as nn

from .. import ops

__all__ = ['av_wrapper']


class _HeadFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, head, x, *params):
        x = x.detach().contiguous().float()
        acts = [x]
        layers = head._linears()
        for i, lin in enumerate(layers):
            acts.append(ops.linear_forward(acts[-1], lin.weight.detach(), lin.bias.detach(), relu=i < len(layers) - 1))
        ctx.head, ctx.acts, ctx.params = head, acts, params
        return acts[-1]

    @staticmethod
    def backward(ctx, dy):
        layers, acts = ctx.head._linears(), ctx.acts
        grads = {}
        d = dy.contiguous().clone()
        for i in reversed(range(len(layers))):
            lin = layers[i]
            relu = i < len(layers) - 1
            dx, dw, db = ops.linear_backward(acts[i], lin.weight.detach(), acts[i + 1] if relu else None, d, relu, need_dx=True)
            grads[lin.weight], grads[lin.bias] = dw, db
            d = dx
        ctx.acts = None
        return (None, d) + tuple(grads.get(p) for p in ctx.params)


class Head(nn.Module):
    """Linear(+ReLU) chain (av_wrapper.py:17-33); nn.Linear children hold the parameters, the math is libavid_b200's."""

    def __init__(self, input_dim, proj_dims):
        super().__init__()
        if not isinstance(proj_dims, list):
            proj_dims = [proj_dims]
        projection = []
        for i, d in enumerate(proj_dims):
            projection += [nn.Linear(input_dim, d)]
            input_dim = d
            if i < len(proj_dims) - 1:
                projection += [nn.ReLU(inplace=True)]
        self.projection = nn.Sequential(*projection)
        self.out_dim = proj_dims[-1]

    def _linears(self):
        return [m for m in self.projection if isinstance(m, nn.Linear)]

    def forward(self, x):
        return _HeadFunction.apply(self, x, *tuple(self.parameters()))


class AV_Wrapper(nn.Module):
    def __init__(self, video_model, audio_model, proj_dim=128):
        super().__init__()
        self.video_model = video_model
        self.audio_model = audio_model
        self.use_linear_proj = proj_dim is not None
        if proj_dim is not None:
            self.video_proj = Head(video_model.out_dim, proj_dim)
            self.audio_proj = Head(audio_model.out_dim, proj_dim)
            self.out_dim = self.video_proj.out_dim
        else:
            self.out_dim = video_model.out_dim

    def forward(self, video, audio):
        video_emb = self.video_model(video)
        video_emb = video_emb.view(video_emb.shape[0], video_emb.shape[1])
        if self.use_linear_proj:
            video_emb = self.video_proj(video_emb)
        audio_emb = self.audio_model(audio)
        audio_emb = audio_emb.view(audio_emb.shape[0], audio_emb.shape[1])
        if self.use_linear_proj:
            audio_emb = self.audio_proj(audio_emb)
        return video_emb, audio_emb


def av_wrapper(video_backbone, video_backbone_args, audio_backbone, audio_backbone_args, proj_dim=128, checkpoint=None):
    """Factory with the reference's signature (av_wrapper.py:64-76): backbones are resolved by name in this package."""
    from .. import models
    assert video_backbone in models.__dict__, 'Unknown model architecture'
    assert audio_backbone in models.__dict__, 'Unknown model architecture'
    video_model = models.__dict__[video_backbone](**video_backbone_args)
    audio_model = models.__dict__[audio_backbone](**audio_backbone_args)
    model = AV_Wrapper(video_model, audio_model, proj_dim=proj_dim)
    if checkpoint is not None:
        ckp = torch.load(checkpoint, map_location='cpu', weights_only=False)
        nn.DataParallel(model).load_state_dict(ckp['model'])   # published checkpoints carry the `module.` prefix
    return model
