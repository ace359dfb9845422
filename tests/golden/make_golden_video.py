#!/usr/bin/env python
"""Generate tests/golden/video_prep.npz by running the UNMODIFIED reference `VideoPrep_MSC_CJ` (datasets/preprocessing.py:15-57,
imported read-only from /root/reference) on seeded random clips.

Runs only in the build container (the GPU box has no /root/reference); the .npz output is committed.
Usage:  python tests/golden/make_golden_video.py

The reference package imports `librosa` and `av` (audio half / decoding, unused here) and `torchvision` (five functional wrappers over
Pillow); none is installed.  Both are injected as stub modules BEFORE the import -- `torchvision.transforms.functional` with the five functions
restated from torchvision 0.5.0 on Pillow (oracle/video.py `tv_*`), `librosa` and `av` empty.  No reference file is touched: the classes that
run -- VideoPrep_MSC_CJ, Compose, RandomResizedCrop (get_params and its RNG order), RandomHorizontalFlip, ColorJitter (get_params, the
shuffle, the per-image loop), ClipToTensor, Normalize -- are the reference's own, on the installed Pillow.
"""
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(1, "/root/reference")

from oracle import video as V  # noqa: E402

tv = types.ModuleType("torchvision")
tv.transforms = types.ModuleType("torchvision.transforms")
fn = types.ModuleType("torchvision.transforms.functional")
fn.resized_crop = V.tv_resized_crop
fn.adjust_brightness = V.tv_adjust_brightness
fn.adjust_contrast = V.tv_adjust_contrast
fn.adjust_saturation = V.tv_adjust_saturation
fn.adjust_hue = V.tv_adjust_hue
tv.transforms.functional = fn
av = types.ModuleType("av")                  # utils/ioutils/av_wrappers.py:11 calls av.logging.set_level(0) at import time
av.logging = types.SimpleNamespace(set_level=lambda level: None)
sys.modules.update({"torchvision": tv, "torchvision.transforms": tv.transforms, "torchvision.transforms.functional": fn,
                    "librosa": types.ModuleType("librosa"), "av": av})

from PIL import Image  # noqa: E402
from datasets.preprocessing import VideoPrep_MSC_CJ  # noqa: E402  (reference)

# (name, frames, height, width, crop, seeds): small clips; one at the reference's 224 crop from a 256 x 340 frame (Kinetics short side 256)
CASES = [("small", 4, 96, 128, 64, (0, 1, 2, 3, 4, 5, 6, 7)),
         ("tall", 3, 150, 100, 56, (11, 12, 13)),
         ("k400", 2, 256, 340, 224, (21,))]


def main():
    out = {}
    for name, t, h, w, crop, seeds in CASES:
        frames = np.random.default_rng(1000 * t + h).integers(0, 256, (t, h, w, 3), dtype=np.uint8)
        # smooth the noise a little so that the jitter sees realistic saturation / hue ranges as well as extremes
        frames[:, : h // 2] = (frames[:, : h // 2].astype(np.int32) // 4 + np.arange(w)[None, None, :, None] * 191 // w).astype(np.uint8)
        out[name + "_frames"] = frames
        prep = VideoPrep_MSC_CJ(crop=(crop, crop), augment=True, num_frames=t, pad_missing=True)
        for seed in seeds:
            random.seed(seed)
            y = prep([Image.fromarray(f) for f in frames])
            out["%s_seed%d" % (name, seed)] = y.numpy()
            assert y.shape == (3, t, crop, crop) and str(y.dtype) == "torch.float32"
    out["cases"] = np.array([(n, str(c)) for n, _, _, _, c, _ in CASES])
    path = os.path.join(HERE, "video_prep.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
