"""Input-side steps on the GPU.  The reference runs them in its CPU loader workers: the log-spectrogram of every
audio clip (datasets/preprocessing.py:158-186, librosa) and the crop / flip / colour-jitter / normalise augmentation of every video
clip (datasets/preprocessing.py:15-57, torchvision on Pillow).  Here clips already on the device are transformed by libavid_b200
(csrc/spectrogram.cu, csrc/video_prep.cu)."""
import ctypes as C
import math
import random

import torch

from .. import _lib
from ..ops import _p, _stream, check


class LogSpectrogram(object):
    """Same constructor arguments as the reference's LogSpectrogram(fps, n_fft, hop_size, normalize); `stats` = (mean, std)
    per output bin replaces the .npz files the reference loads from datasets/assets when normalize=True.
    __call__(sig (B, num_samples) or (B, 1, num_samples) CUDA fp32, duration) -> (B, 1, frames, n_fft // 2 + 1)."""

    def __init__(self, fps, n_fft=512, hop_size=0.005, normalize=False, stats=None, top_db=100.0):
        self.inp_fps, self.n_fft, self.hop_size, self.rate = fps, n_fft, hop_size, 1. / hop_size
        self.normalize, self.top_db = normalize, top_db
        if normalize and stats is None:
            raise ValueError("normalize=True needs stats=(mean, std) with n_fft // 2 + 1 entries each")
        self.mean, self.std = (None, None) if not normalize else [torch.as_tensor(t, dtype=torch.float32) for t in stats]

    def __call__(self, sig, sr=None, duration=None):
        if not sig.is_cuda:
            raise RuntimeError("avid_cma_b200 LogSpectrogram runs on CUDA tensors only (there is no CPU path)")
        sr = self.inp_fps if sr is None else sr
        sig = sig.reshape(sig.shape[0], -1).contiguous().float()
        B, L = sig.shape
        hop = int(self.hop_size * sr)
        frames = 1 + L // hop
        if duration is not None:
            frames = min(frames, int(duration * self.rate))
        bins = self.n_fft // 2 + 1
        out = torch.empty(B, 1, frames, bins, dtype=torch.float32, device=sig.device)
        L_ = _lib.lib()
        ws = torch.empty(int(L_.avid_log_spectrogram_workspace_bytes(B)), dtype=torch.uint8, device=sig.device)
        mean = std = None
        if self.normalize:
            mean, std = self.mean.to(sig.device), self.std.to(sig.device)
            if mean.numel() != bins or std.numel() != bins:
                raise ValueError("stats must have %d entries" % bins)
        check(L_.avid_log_spectrogram(_p(sig), B, L, 2 * self.n_fft, hop, frames, float(-1.0 if self.top_db is None else self.top_db),
                                      _p(mean, optional=True), _p(std, optional=True), _p(out), _p(ws, torch.uint8), ws.numel(), _stream()))
        return out, self.rate


class VideoPrep_MSC_CJ(object):
    """Same constructor arguments as the reference's VideoPrep_MSC_CJ (datasets/preprocessing.py:15-57): multi-scale crop
    (RandomResizedCrop, scale = (min_area, 1)), random horizontal flip, colour jitter (brightness, contrast, saturation, hue),
    ClipToTensor, Normalize(ImageNet mean / std).

    __call__(frames): frames = (T, H, W, 3) uint8 CUDA tensor (the decoded clip; the reference passes a list of PIL images) ->
    (3, T, crop_h, crop_w) float32, bit-identical to the reference on Pillow 12 for the same `random` state: the random decisions are
    drawn on the host with the `random` module in the reference's order (RandomResizedCrop.get_params, RandomHorizontalFlip,
    ColorJitter.get_params, the shuffle of the op order), the pixels are produced by `avid_video_prep`.
    `draw(width, height)` returns those decisions, `apply(frames, params)` runs the kernels for given ones."""

    KINDS = {'brightness': 0, 'saturation': 1, 'hue': 2, 'contrast': 3}

    def __init__(self, crop=(224, 224), color=(0.4, 0.4, 0.4, 0.2), min_area=0.08, augment=True, normalize=True, totensor=True,
                 num_frames=8, pad_missing=False):
        if normalize:
            assert totensor
        if not augment:
            raise NotImplementedError("augment=False (the evaluation transform: Resize + CenterCrop) is not on the training hot path")
        if not totensor:
            raise NotImplementedError("the device path returns tensors (totensor=True)")
        self.crop = tuple(crop) if isinstance(crop, (tuple, list)) else (crop, crop)
        self.color, self.min_area = tuple(color), min_area
        self.augment, self.normalize = augment, normalize
        self.num_frames, self.pad_missing = num_frames, pad_missing
        self.mean, self.std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)

    # ---- the random decisions, in the reference's call order (utils/videotransforms/video_transforms.py) ----
    def draw(self, width, height):
        scale, ratio = (self.min_area, 1.), (3. / 4., 4. / 3.)
        area = width * height
        box = None
        for _ in range(10):                                     # RandomResizedCrop.get_params, :330-371
            target_area = random.uniform(*scale) * area
            aspect_ratio = math.exp(random.uniform(math.log(ratio[0]), math.log(ratio[1])))
            w = int(round(math.sqrt(target_area * aspect_ratio)))
            h = int(round(math.sqrt(target_area / aspect_ratio)))
            if w <= width and h <= height:
                i = random.randint(0, height - h)
                j = random.randint(0, width - w)
                box = (i, j, h, w)
                break
        if box is None:                                         # fallback to a central crop
            in_ratio = width / height
            if in_ratio < min(ratio):
                w = width
                h = int(round(w / min(ratio)))
            elif in_ratio > max(ratio):
                h = height
                w = int(round(h * max(ratio)))
            else:
                w, h = width, height
            box = ((height - h) // 2, (width - w) // 2, h, w)
        flip = random.random() < 0.5                            # RandomHorizontalFlip, :86
        brightness, contrast, saturation, hue = self.color      # ColorJitter.get_params, :413-436
        b = random.uniform(max(0, 1 - brightness), 1 + brightness) if brightness > 0 else None
        c = random.uniform(max(0, 1 - contrast), 1 + contrast) if contrast > 0 else None
        s = random.uniform(max(0, 1 - saturation), 1 + saturation) if saturation > 0 else None
        h_ = random.uniform(-hue, hue) if hue > 0 else None
        ops = [(n, f) for n, f in (('brightness', b), ('saturation', s), ('hue', h_), ('contrast', c)) if f is not None]   # :449-457
        random.shuffle(ops)                                     # :458
        return dict(crop=box, flip=flip, ops=ops)

    def _fill(self, frames, params):
        if not frames.is_cuda:
            raise RuntimeError("avid_cma_b200 VideoPrep_MSC_CJ runs on CUDA tensors only (there is no CPU path)")
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[3] != 3:
            raise ValueError("frames must be a (T, H, W, 3) uint8 tensor")
        T, H, W, _ = frames.shape
        ops = params['ops']
        kinds, factors, hue_shift = [0, 0, 0, 0], [0., 0., 0., 0.], 0
        for k, (name, f) in enumerate(ops):
            kinds[k], factors[k] = self.KINDS[name], f
            if name == 'hue':
                if not (-0.5 <= f <= 0.5):
                    raise ValueError('hue_factor is not in [-0.5, 0.5].')
                hue_shift = int(f * 255) & 0xFF                 # np.uint8(hue_factor * 255) of torchvision's adjust_hue
        i, j, h, w = params['crop']
        return _lib.VideoPrep(T, H, W, i, j, h, w, self.crop[0], self.crop[1], int(bool(params['flip'])), len(ops), self._i4(*kinds), self._f4(*factors),
                              hue_shift, int(self.normalize), self._f3(*self.mean), self._f3(*self.std))

    _i4, _f4, _f3 = C.c_int32 * 4, C.c_float * 4, C.c_float * 3

    def apply(self, frames, params):
        """One clip (T, H, W, 3) uint8 -> (3, T, crop_h, crop_w) float32 with the given decisions."""
        return self.apply_batch([frames], [params])[0]

    def plan_batch(self, clips, params):
        """Everything of a batched call that happens on the host: parameter structs, output buffer, workspace, pointer arrays.
        Returns (run, outs): `run()` enqueues the kernels (one library call: 4 launches per 16 clips) and returns `outs`, a list of
        (3, T, crop_h, crop_w) float32 tensors -- slices of ONE (B, 3, T, h, w) buffer when the frame counts agree (outs.base)."""
        clips = [c if c.is_contiguous() else c.contiguous() for c in clips]
        n = len(clips)
        P = (_lib.VideoPrep * n)(*[self._fill(c, q) for c, q in zip(clips, params)])
        dev = clips[0].device
        L_ = _lib.lib()
        need = int(L_.avid_video_prep_batch_workspace_bytes(P, n))       # 0 on invalid parameters: the call reports which
        ws = torch.empty(max(need, 1), dtype=torch.uint8, device=dev)
        T = clips[0].shape[0]
        if all(c.shape[0] == T for c in clips):
            base = torch.empty(n, 3, T, self.crop[0], self.crop[1], dtype=torch.float32, device=dev)
            outs = list(base.unbind(0))
            p0, step = base.data_ptr(), base.stride(0) * 4
            op = (C.c_void_p * n)(*[p0 + k * step for k in range(n)])
        else:
            outs = [torch.empty(3, c.shape[0], self.crop[0], self.crop[1], dtype=torch.float32, device=dev) for c in clips]
            op = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
        fp = (C.c_void_p * n)(*[c.data_ptr() for c in clips])
        wp, wn = C.c_void_p(ws.data_ptr()), ws.numel()

        def run(_keep=(clips, ws, P, fp, op)):
            check(L_.avid_video_prep_batch(fp, P, n, op, wp, wn, _stream()))
            return outs
        return run, outs

    def apply_batch(self, clips, params):
        """`clips`: list of (T, H, W, 3) uint8 CUDA tensors (sizes may differ), `params`: one dict of decisions per clip."""
        return self.plan_batch(clips, params)[0]()

    def __call__(self, frames):
        if isinstance(frames, (list, tuple)) or frames.dim() == 5:      # a loader batch: decisions drawn clip by clip, one library call
            clips = list(frames)
            outs = self.apply_batch(clips, [self.draw(c.shape[2], c.shape[1]) for c in clips])
            return torch.stack([self._pad(o) for o in outs])
        return self._pad(self.apply(frames, self.draw(frames.shape[2], frames.shape[1])))

    def _pad(self, out):
        if self.pad_missing:                                    # preprocessing.py:49-56
            while True:
                n_missing = self.num_frames - out.shape[1]
                if n_missing > 0:
                    out = torch.cat((out, out[:, :n_missing]), 1)
                else:
                    break
        return out
