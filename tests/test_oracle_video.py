"""CPU: the video input-step oracle (oracle/video.py) is pinned
  * against Pillow, the library the reference's transforms execute: the numpy restatement of the resampler, the ImageEnhance blends,
    L and RGB <-> HSV (ALL 2^24 colours) is bit-identical;
  * against tests/golden/video_prep.npz, the outputs of the UNMODIFIED reference VideoPrep_MSC_CJ (tests/golden/make_golden_video.py):
    drawing the parameters with the same `random` seed reproduces them exactly (RNG call order, shuffle, ClipToTensor, Normalize)."""
import os
import random

import numpy as np
import pytest

from oracle import video as V

PIL = pytest.importorskip("PIL")
from PIL import Image, ImageEnhance  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "video_prep.npz")


def test_luma_and_hsv_match_pillow_for_every_colour():
    allc = np.stack(np.meshgrid(np.arange(256), np.arange(256), np.arange(256), indexing="ij"), -1).reshape(4096, 4096, 3).astype(np.uint8)
    im = Image.fromarray(allc)
    assert np.array_equal(np.array(im.convert("L")), V.luma_np(allc))
    assert np.array_equal(np.array(im.convert("HSV")), V.rgb2hsv_np(allc))
    assert np.array_equal(np.array(Image.fromarray(allc, "HSV").convert("RGB")), V.hsv2rgb_np(allc))


@pytest.mark.parametrize("f", [0.0, 1.0, 0.6, 0.6123, 1.4, 1.3999, 0.999, 1.001, 0.123456])
def test_enhance_blends_match_pillow(f):
    img = np.random.default_rng(3).integers(0, 256, (64, 80, 3), dtype=np.uint8)
    img[:32] //= 3
    pim = Image.fromarray(img)
    assert np.array_equal(np.array(ImageEnhance.Brightness(pim).enhance(f)), V.brightness_np(img, f))
    assert np.array_equal(np.array(ImageEnhance.Color(pim).enhance(f)), V.saturation_np(img, f))
    assert np.array_equal(np.array(ImageEnhance.Contrast(pim).enhance(f)), V.contrast_np(img, f))


@pytest.mark.parametrize("f", [-0.2, -0.1, -0.003, 0.0, 0.05, 0.1999, 0.2])
def test_hue_matches_pillow_and_numpy_uint8_cast(f):
    img = np.random.default_rng(4).integers(0, 256, (40, 50, 3), dtype=np.uint8)
    assert np.array_equal(np.array(V.tv_adjust_hue(Image.fromarray(img), f)), V.hue_np(img, f))
    with np.errstate(all="ignore"):
        assert int(np.array(f * 255).astype(np.uint8)) == V.hue_shift_u8(f)        # what np.uint8(hue_factor * 255) does in F.adjust_hue


@pytest.mark.parametrize("hw,box,size", [((120, 160), (3, 5, 100, 140), (64, 64)), ((240, 320), (10, 20, 200, 250), (224, 224)),
                                         ((64, 64), (0, 0, 30, 20), (48, 56)), ((100, 90), (5, 5, 90, 80), (90, 40)),
                                         ((50, 50), (1, 2, 7, 9), (32, 32)), ((128, 171), (0, 20, 128, 128), (112, 112))])
def test_resized_crop_matches_pillow(hw, box, size):
    frames = np.random.default_rng(5).integers(0, 256, (2,) + hw + (3,), dtype=np.uint8)
    i, j, h, w = box
    ref = np.stack([np.array(V.tv_resized_crop(Image.fromarray(f), i, j, h, w, size, Image.BILINEAR)) for f in frames])
    assert np.array_equal(ref, V.resized_crop_np(frames, i, j, h, w, size))


def test_pipeline_reproduces_the_reference_goldens():
    g = np.load(GOLD)
    n = 0
    for name, crop in g["cases"]:
        crop = int(crop)
        frames = g[name + "_frames"]
        for key in [k for k in g.files if k.startswith(name + "_seed")]:
            random.seed(int(key.split("seed")[1]))
            params = V.draw_params(frames.shape[2], frames.shape[1])
            want = g[key]
            assert np.array_equal(V.video_prep_pil(frames, params, crop=(crop, crop)), want), key
            assert np.array_equal(V.video_prep_np(frames, params, crop=(crop, crop)), want), key
            n += 1
    assert n >= 10
    # the goldens exercise both flip states and every position of the contrast op in the shuffled order
    flips, pos = set(), set()
    for name, crop in g["cases"]:
        frames = g[name + "_frames"]
        for key in [k for k in g.files if k.startswith(name + "_seed")]:
            random.seed(int(key.split("seed")[1]))
            p = V.draw_params(frames.shape[2], frames.shape[1])
            flips.add(p["flip"])
            pos.add([o[0] for o in p["ops"]].index("contrast"))
    assert flips == {True, False} and len(pos) >= 3
