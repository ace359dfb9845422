"""Input-side step on the GPU: the reference computes the log-spectrogram of every audio clip in its CPU
loader workers (datasets/preprocessing.py:158-186, librosa); here a batch of mono clips already on the device is
transformed by libavid_b200 (csrc/spectrogram.cu)."""
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
