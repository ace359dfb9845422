"""GPU log-spectrogram (csrc/spectrogram.cu) against the numpy restatement of the reference's LogSpectrogram
(oracle/audio.py; librosa is not available, so this row is pinned to the restatement only)."""
import numpy as np
import pytest
import torch

from oracle import audio as oa

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _clips(B, L, seed):
    g = np.random.default_rng(seed)
    t = np.arange(L) / 24000.0
    sig = 0.1 * g.standard_normal((B, L))
    for b in range(B):
        sig[b] += 0.5 * np.sin(2 * np.pi * (220.0 * (b + 1)) * t) + 0.2 * np.sin(2 * np.pi * 3150.0 * t + b)
    return sig.astype(np.float32)


@pytest.mark.parametrize("n_fft,hop_size,duration,seconds", [(512, 0.01, 2.0, 2.0), (256, 0.01, 1.0, 1.3), (512, 0.005, None, 0.5)])
def test_log_spectrogram_vs_numpy_oracle(n_fft, hop_size, duration, seconds):
    from avid_cma_b200.datasets.gpu_preprocessing import LogSpectrogram
    sr, B = 24000, 3
    sig = _clips(B, int(seconds * sr), seed=n_fft)
    bins = n_fft // 2 + 1
    g = np.random.default_rng(1)
    mean, std = (-30 + 5 * g.standard_normal(bins)).astype(np.float32), (15 + 3 * g.random(bins)).astype(np.float32)
    for stats in (None, (mean, std)):
        op = LogSpectrogram(sr, n_fft=n_fft, hop_size=hop_size, normalize=stats is not None, stats=stats)
        got, rate = op(torch.from_numpy(sig).to(DEV).unsqueeze(1), sr, duration)
        want = np.stack([oa.log_spectrogram(sig[b], sr, n_fft, hop_size, duration, *(stats or (None, None))) for b in range(B)])
        assert rate == 1.0 / hop_size and tuple(got.shape) == want.shape
        scale = 1.0 if stats is None else 1.0 / 15.0
        err = np.abs(got.cpu().numpy().astype(np.float64) - want)
        # fp32 FFT: bins 60+ dB below the clip maximum carry rounding noise of the strong bins; everything else is tight
        floor = want.reshape(B, -1).max(1)[:, None, None, None] - (60.0 if stats is None else 1e9)
        strong = want > floor if stats is None else np.ones_like(want, dtype=bool)
        assert err[strong].max() < (2e-3 if stats is None else 0.2) and np.median(err) < 2e-4 * max(scale, 1.0)


def test_log_spectrogram_shape_of_the_training_config():
    """Kinetics config: 2 s at 24 kHz, n_fft 512 (1024-point frames), 100 frames / s -> (B, 1, 200, 257) as the audio tower expects."""
    from avid_cma_b200.datasets.gpu_preprocessing import LogSpectrogram
    sig = torch.from_numpy(_clips(4, 48000, seed=3)).to(DEV)
    out, _ = LogSpectrogram(24000, n_fft=512, hop_size=0.01)(sig, 24000, 2.0)
    assert tuple(out.shape) == (4, 1, 200, 257) and bool(torch.isfinite(out).all())
    assert float(out.amax(dim=(1, 2, 3)).min()) - float(out.amin()) <= 100.0 + 1e-3      # top_db floor per clip
    with pytest.raises(RuntimeError):
        LogSpectrogram(24000)(torch.zeros(1, 48000))
