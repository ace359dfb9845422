"""GPU video input step (csrc/video_prep.cu behind datasets.gpu_preprocessing.VideoPrep_MSC_CJ) -- BIT-EXACT against
  * tests/golden/video_prep.npz: outputs of the UNMODIFIED reference VideoPrep_MSC_CJ (tests/golden/make_golden_video.py), same `random` seeds;
  * the oracle (oracle/video.py, pinned to Pillow) on fresh seeded clips: every op alone, every position of the contrast op, up- and
    down-scaling crops, ragged extents, flip on / off, normalize off;
  * Pillow itself where it is importable on the box.
uint8 / index work: the bar is exact equality of every output value."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import video as V

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "video_prep.npz")


def _prep(crop, **kw):
    from avid_cma_b200.datasets.gpu_preprocessing import VideoPrep_MSC_CJ
    return VideoPrep_MSC_CJ(crop=crop, **kw)


def _clip(t, h, w, seed):
    g = np.random.default_rng(seed)
    f = g.integers(0, 256, (t, h, w, 3), dtype=np.uint8)
    f[:, : h // 2] = (f[:, : h // 2].astype(np.int32) // 5 + (np.arange(w)[None, None, :, None] * 200) // w).astype(np.uint8)      # smooth half
    f[:, :, : w // 3, 1] = f[:, :, : w // 3, 0]                                                                                      # low saturation
    return f


def test_reproduces_the_reference_goldens_bit_for_bit():
    g = np.load(GOLD)
    n = 0
    for name, crop in g["cases"]:
        crop = int(crop)
        frames = g[name + "_frames"]
        prep = _prep((crop, crop), num_frames=frames.shape[0], pad_missing=True)
        dev_frames = torch.from_numpy(frames).to(DEV)
        for key in [k for k in g.files if k.startswith(name + "_seed")]:
            random.seed(int(key.split("seed")[1]))
            got = prep(dev_frames).cpu().numpy()
            assert got.dtype == np.float32 and got.shape == g[key].shape
            assert np.array_equal(got, g[key]), (key, int((got != g[key]).sum()))
            n += 1
    assert n >= 10


OPS = [[], [("brightness", 0.7)], [("brightness", 1.3)], [("saturation", 0.65)], [("saturation", 1.38)], [("hue", -0.17)], [("hue", 0.2)],
       [("contrast", 0.61)], [("contrast", 1.4)],
       [("contrast", 1.21), ("hue", 0.05), ("brightness", 0.9), ("saturation", 1.1)],
       [("hue", -0.2), ("contrast", 0.8), ("saturation", 0.7), ("brightness", 1.39)],
       [("saturation", 1.25), ("brightness", 0.66), ("contrast", 1.33), ("hue", 0.11)],
       [("brightness", 1.0), ("saturation", 0.0), ("hue", 0.0), ("contrast", 1.0)]]


@pytest.mark.parametrize("k", range(len(OPS)))
def test_every_op_and_order_vs_oracle(k):
    frames = _clip(3, 70, 90, seed=k)
    params = dict(crop=(4, 7, 50, 61), flip=bool(k & 1), ops=OPS[k])
    got = _prep((40, 48)).apply(torch.from_numpy(frames).to(DEV), params).cpu().numpy()
    want = V.video_prep_np(frames, params, crop=(40, 48))
    assert np.array_equal(got, want), int((got != want).sum())


@pytest.mark.parametrize("hw,box,size", [((120, 160), (3, 5, 100, 140), (64, 64)),      # down-scaling both axes
                                         ((64, 64), (0, 0, 30, 20), (48, 56)),           # up-scaling
                                         ((100, 90), (5, 5, 90, 80), (90, 40)),          # height unchanged (Pillow skips that pass)
                                         ((50, 50), (1, 2, 7, 9), (32, 32)),             # tiny crop
                                         ((256, 340), (0, 0, 256, 340), (224, 224)),     # the training geometry, whole frame
                                         ((360, 640), (13, 100, 347, 500), (112, 112))])  # stronger down-scaling (5 taps and more)
def test_resampler_geometries_vs_oracle_and_pillow(hw, box, size):
    frames = _clip(2, hw[0], hw[1], seed=hw[0])
    params = dict(crop=box, flip=False, ops=[])
    prep = _prep(size, normalize=False)
    got = prep.apply(torch.from_numpy(frames).to(DEV), params).cpu().numpy()
    want = V.video_prep_np(frames, params, crop=size, normalize=False)
    assert np.array_equal(got, want), int((got != want).sum())
    try:
        import PIL  # noqa: F401
    except ImportError:
        return
    assert np.array_equal(got, V.video_prep_pil(frames, params, crop=size, normalize=False))


def test_random_draws_full_pipeline_at_training_size():
    """BASELINE config-2 input geometry (8 frames, 224 x 224 crop of a 256 x 340 frame): 6 random draws against the oracle."""
    frames = _clip(8, 256, 340, seed=99)
    prep = _prep((224, 224))
    dev_frames = torch.from_numpy(frames).to(DEV)
    for seed in range(6):
        random.seed(1000 + seed)
        params = prep.draw(340, 256)
        random.seed(1000 + seed)
        assert params == V.draw_params(340, 256)          # the product's host logic draws what the oracle (= the reference) draws
        got = prep.apply(dev_frames, params).cpu().numpy()
        want = V.video_prep_np(frames, params, crop=(224, 224))
        assert got.shape == (3, 8, 224, 224) and np.array_equal(got, want), (seed, int((got != want).sum()))


def test_batched_call_equals_clip_by_clip():
    """One library call for a loader batch (ragged clip sizes, 19 clips = two launch groups) == clip-by-clip calls, same `random` stream."""
    prep = _prep((32, 40))
    clips = [torch.from_numpy(_clip(2 + k % 3, 40 + 3 * k, 50 + 2 * k, seed=k)).to(DEV) for k in range(19)]
    random.seed(5)
    params = [prep.draw(c.shape[2], c.shape[1]) for c in clips]
    one = [prep.apply(c, q) for c, q in zip(clips, params)]
    many = prep.apply_batch(clips, params)
    assert all(torch.equal(a, b) for a, b in zip(one, many))
    same = torch.stack([torch.from_numpy(_clip(4, 48, 64, seed=100 + k)) for k in range(5)]).to(DEV)      # (B, T, H, W, 3)
    random.seed(6)
    out = prep(same)
    random.seed(6)
    assert tuple(out.shape) == (5, 3, 4, 32, 40) and all(torch.equal(out[k], prep(same[k])) for k in range(5))


def test_many_frames_small_output_with_contrast():
    """More frames than output columns: every frame's luma sum must start at zero (the sums are cleared by the weight kernel)."""
    frames = _clip(40, 24, 20, seed=11)
    params = dict(crop=(2, 1, 20, 16), flip=True, ops=[("brightness", 1.2), ("contrast", 0.7), ("hue", 0.1)])
    prep = _prep((8, 6))
    dev_frames = torch.from_numpy(frames).to(DEV)
    want = V.video_prep_np(frames, params, crop=(8, 6))
    for _ in range(2):                                   # twice: the second call reuses a workspace with stale sums
        assert np.array_equal(prep.apply(dev_frames, params).cpu().numpy(), want)


def test_arguments_and_edges():
    prep = _prep((16, 16), num_frames=5, pad_missing=True)
    frames = torch.from_numpy(_clip(2, 20, 24, seed=5)).to(DEV)
    random.seed(3)
    out = prep(frames)
    assert tuple(out.shape) == (3, 5, 16, 16)                              # pad_missing repeats frames like preprocessing.py:49-56
    assert torch.equal(out[:, 2:4], out[:, 0:2]) and torch.equal(out[:, 4], out[:, 0])
    with pytest.raises(RuntimeError):
        prep.apply(frames.cpu(), dict(crop=(0, 0, 20, 24), flip=False, ops=[]))
    with pytest.raises(RuntimeError):                                      # crop box outside the frame -> EINVAL from the library
        prep.apply(frames, dict(crop=(0, 0, 21, 24), flip=False, ops=[]))
    with pytest.raises(ValueError):
        prep.apply(frames, dict(crop=(0, 0, 20, 24), flip=False, ops=[("hue", 0.6)]))
    with pytest.raises(NotImplementedError):
        _prep((16, 16), augment=False)
