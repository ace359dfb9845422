"""Two-tower wrapper + projection heads (reference: models/av_wrapper.py)."""
import os

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


def _mlp(widths):
    """[Linear, ReLU, Linear, ..., Linear] over consecutive width pairs: the Linear layers sit at the even positions of the
    Sequential ('projection.0', 'projection.2', ...), which is the parameter naming of the reference's Head (av_wrapper.py:17-33)."""
    mods = []
    for k, (fan_in, fan_out) in enumerate(zip(widths[:-1], widths[1:])):
        if k:
            mods.append(nn.ReLU(inplace=True))
        mods.append(nn.Linear(fan_in, fan_out))
    return nn.Sequential(*mods)


class Head(nn.Module):
    """Projection head: Linear(+ReLU) chain; the nn.Linear children only hold the parameters, the math is libavid_b200's
    (one autograd.Function for the whole chain)."""

    def __init__(self, input_dim, proj_dims):
        super().__init__()
        widths = [input_dim] + (list(proj_dims) if isinstance(proj_dims, (list, tuple)) else [proj_dims])
        self.projection = _mlp(widths)
        self.out_dim = widths[-1]

    def _linears(self):
        return [m for m in self.projection if isinstance(m, nn.Linear)]

    def forward(self, x):
        return _HeadFunction.apply(self, x, *tuple(self.parameters()))


class AV_Wrapper(nn.Module):
    """forward(video, audio) -> (video_emb, audio_emb), both (B, out_dim) (av_wrapper.py:36-61).  proj_dim=None: the pooled
    512-d tower outputs are returned without heads."""

    def __init__(self, video_model, audio_model, proj_dim=128):
        super().__init__()
        self.video_model, self.audio_model = video_model, audio_model
        self.use_linear_proj = proj_dim is not None
        self.out_dim = video_model.out_dim
        if self.use_linear_proj:
            self.video_proj, self.audio_proj = Head(video_model.out_dim, proj_dim), Head(audio_model.out_dim, proj_dim)
            self.out_dim = self.video_proj.out_dim

    def _embed(self, tower, head_name, x):
        pooled = tower(x)
        emb = pooled.view(pooled.shape[0], pooled.shape[1])          # (B, C, 1, 1[, 1]) -> (B, C)
        return getattr(self, head_name)(emb) if self.use_linear_proj else emb

    def forward(self, video, audio):
        if os.environ.get('AVID_TOWER_STREAMS', '1') == '1' and video.is_cuda:
            # the two towers are independent until the criterion: the audio tower (11 % of the FLOPs) runs on a side stream, so
            # its bandwidth-bound passes overlap the video tower's tensor-bound ones; autograd replays each tower's backward on
            # the stream of its forward
            cur = torch.cuda.current_stream()
            side = self.__dict__.get('_side_stream')
            if side is None:
                side = self.__dict__['_side_stream'] = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                audio_emb = self._embed(self.audio_model, 'audio_proj', audio)
            video_emb = self._embed(self.video_model, 'video_proj', video)
            cur.wait_stream(side)
            audio_emb.record_stream(cur)
            return video_emb, audio_emb
        return self._embed(self.video_model, 'video_proj', video), self._embed(self.audio_model, 'audio_proj', audio)


def av_wrapper(video_backbone, video_backbone_args, audio_backbone, audio_backbone_args, proj_dim=128, checkpoint=None):
    """Factory with the reference's signature (av_wrapper.py:64-76): backbones are resolved by name in this package; a
    checkpoint written under DataParallel / DistributedDataParallel (`module.`-prefixed keys) is loaded the same way."""
    from .. import models
    towers = []
    for name, kwargs in ((video_backbone, video_backbone_args), (audio_backbone, audio_backbone_args)):
        if name not in models.__dict__:
            raise AssertionError('Unknown model architecture')
        towers.append(models.__dict__[name](**kwargs))
    model = AV_Wrapper(towers[0], towers[1], proj_dim=proj_dim)
    if checkpoint is not None:
        state = torch.load(checkpoint, map_location='cpu', weights_only=False)['model']
        nn.DataParallel(model).load_state_dict(state)
    return model
