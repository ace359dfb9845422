"""CPU oracle of the audio input step (TEST INFRASTRUCTURE: imported by tests/ only).

Restates LogSpectrogram.__call__ of the reference (datasets/preprocessing.py:158-186) in numpy fp64.  The arithmetic of
that function lives in librosa (`librosa.stft`, `librosa.core.power_to_db`), which is neither in /root/reference nor
installed here, and the reference does not pin its version (conda-spec-list.txt lists no librosa): PARITY UNPINNED for this
row.  The restatement follows the librosa 0.7-0.9 defaults the 2020 code base was written against: stft(n_fft, hop_length,
win_length = n_fft, window = 'hann' (periodic), center = True, pad_mode = 'reflect'); power_to_db(S, ref = 1.0,
amin = 1e-10, top_db)."""
import numpy as np


def stft_power(sig, n_fft, hop):
    """|STFT|^2, shape (n_fft // 2 + 1, 1 + len(sig) // hop) like librosa.stft with centred, reflect-padded frames."""
    sig = np.asarray(sig, dtype=np.float64)
    window = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n_fft) / n_fft)          # scipy.signal.get_window('hann', n_fft, fftbins=True)
    padded = np.pad(sig, n_fft // 2, mode="reflect")
    n_frames = 1 + (len(padded) - n_fft) // hop
    frames = np.stack([padded[i * hop:i * hop + n_fft] * window for i in range(n_frames)], axis=1)
    return np.abs(np.fft.rfft(frames, axis=0)) ** 2


def power_to_db(S, amin=1e-10, top_db=100.0):
    db = 10.0 * np.log10(np.maximum(amin, S)) - 10.0 * np.log10(np.maximum(amin, 1.0))
    if top_db is not None:
        db = np.maximum(db, db.max() - top_db)
    return db


def log_spectrogram(sig, sr, n_fft=512, hop_size=0.01, duration=None, mean=None, std=None, top_db=100.0):
    """preprocessing.py:168-186 for one mono clip: returns (1, frames, n_fft // 2 + 1) float64."""
    hop = int(hop_size * sr)
    spect = stft_power(sig, n_fft * 2, hop)
    spect = np.concatenate([spect[:1], spect[1:].reshape(n_fft // 2, 2, -1).mean(1)], 0)
    if duration is not None:
        spect = spect[:, :int(duration * (1.0 / hop_size))]
    spect = power_to_db(spect, top_db=top_db)
    if mean is not None:
        spect = (spect - np.asarray(mean, dtype=np.float64)[:, None]) / (np.asarray(std, dtype=np.float64)[:, None] + 1e-5)
    return spect.T[None]
