"""Synthetic audio-visual clips with the sample contract of the reference's VideoDataset (datasets/video_db.py:219-265 in
'clip' mode with return_index=True): {'frames': (3, T, H, W) float32, 'audio': (1, F, S) float32, 'index': int}.

`index` is the instance id the memory bank is addressed with; with clips_per_video > 1 the same instance appears several
times per epoch (video_db.py:98: index % num_samples), which is reproduced here."""
import torch
from torch.utils import data


class SyntheticAV(data.Dataset):
    def __init__(self, num_samples, num_frames=8, crop_size=224, spectrogram=(200, 257), clips_per_video=1, seed=0):
        self.num_samples, self.clips_per_video = int(num_samples), int(clips_per_video)
        self.num_frames, self.crop_size, self.spectrogram, self.seed = int(num_frames), int(crop_size), tuple(spectrogram), int(seed)

    def __len__(self):
        return self.num_samples * self.clips_per_video

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 1000003 + i)
        return {'frames': torch.randn(3, self.num_frames, self.crop_size, self.crop_size, generator=g),
                'audio': torch.randn(1, self.spectrogram[0], self.spectrogram[1], generator=g),
                'index': i % self.num_samples}

    def __repr__(self):
        return ('SyntheticAV(num_samples={}, clips_per_video={}, frames=3x{}x{}x{}, audio=1x{}x{})'
                .format(self.num_samples, self.clips_per_video, self.num_frames, self.crop_size, self.crop_size, *self.spectrogram))
