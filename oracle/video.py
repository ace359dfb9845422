"""CPU oracle of the video input step (TEST INFRASTRUCTURE: imported by tests/ and tests/golden/make_golden_video.py only).

Restates `VideoPrep_MSC_CJ.__call__` of the reference (datasets/preprocessing.py:15-57) for one clip of uint8 RGB frames:
RandomResizedCrop -> RandomHorizontalFlip -> ColorJitter -> ClipToTensor -> Normalize
(utils/videotransforms/video_transforms.py:73-98,303-391,393-476, volume_transforms.py:14-70, tensor_transforms.py:13-38).

Where the arithmetic lives.  The reference's transforms call torchvision.transforms.functional on PIL images, and those five
functions (`resized_crop`, `adjust_brightness`, `adjust_contrast`, `adjust_saturation`, `adjust_hue`) are thin wrappers over
Pillow: `Image.crop` + `Image.resize(BILINEAR)`, `ImageEnhance.{Brightness,Contrast,Color}.enhance`, and an HSV round trip
with a uint8 hue shift.  Pillow IS installed here (12.2) and is what the numpy restatements below are pinned against, bit for
bit (tests/test_oracle_video.py: random images for the resampler and the blends, ALL 2^24 colours for RGB <-> HSV and L).
torchvision is absent (the reference's conda spec pins torchvision 0.5/0.6 era semantics): its five wrappers are restated from
the published source of `torchvision/transforms/functional.py` (v0.5.0) in `tv_*` below; everything around them -- the RNG call
order of the `get_params` functions, the shuffle of the jitter order, ClipToTensor, Normalize -- is pinned by running the
UNMODIFIED reference classes with those wrappers injected as a stub `torchvision` module (tests/golden/make_golden_video.py ->
tests/golden/video_prep.npz).

Two layers:
  * `*_pil`  : the step done by calling Pillow itself (what the reference executes);
  * `*_np`   : the same step as integer / float32 / float64 numpy arithmetic following Pillow's C sources
               (src/libImaging/Resample.c, Blend.c, Convert.c) -- the form the CUDA kernels implement (csrc/video_prep.cu).
"""
import math
import random

import numpy as np

PRECISION_BITS = 32 - 8 - 2          # Resample.c: fixed-point precision of the 8-bit-per-channel convolution


# ---------------------------------------------------------------------------------------------------------------------------
# parameter drawing: the `random` calls of the reference in their order
# ---------------------------------------------------------------------------------------------------------------------------
def draw_crop(width, height, scale=(0.08, 1.0), ratio=(3. / 4., 4. / 3.), rng=random):
    """RandomResizedCrop.get_params (video_transforms.py:330-371) -> (i, j, h, w) = (top, left, height, width)."""
    area = width * height
    for _ in range(10):
        target_area = rng.uniform(*scale) * area
        log_ratio = (math.log(ratio[0]), math.log(ratio[1]))
        aspect_ratio = math.exp(rng.uniform(*log_ratio))
        w = int(round(math.sqrt(target_area * aspect_ratio)))
        h = int(round(math.sqrt(target_area / aspect_ratio)))
        if w <= width and h <= height:
            i = rng.randint(0, height - h)
            j = rng.randint(0, width - w)
            return i, j, h, w
    in_ratio = width / height
    if in_ratio < min(ratio):
        w = width
        h = int(round(w / min(ratio)))
    elif in_ratio > max(ratio):
        h = height
        w = int(round(h * max(ratio)))
    else:
        w, h = width, height
    return (height - h) // 2, (width - w) // 2, h, w


def draw_flip(rng=random):
    """RandomHorizontalFlip.__call__ (video_transforms.py:86): flip iff random() < 0.5."""
    return rng.random() < 0.5


def draw_jitter(brightness=0.4, contrast=0.4, saturation=0.4, hue=0.2, rng=random):
    """ColorJitter.get_params + the shuffle of the op list (video_transforms.py:413-463).
    Returns the ops in application order: [('brightness', f), ('saturation', f), ('hue', f), ('contrast', f)] shuffled."""
    b = rng.uniform(max(0, 1 - brightness), 1 + brightness) if brightness > 0 else None
    c = rng.uniform(max(0, 1 - contrast), 1 + contrast) if contrast > 0 else None
    s = rng.uniform(max(0, 1 - saturation), 1 + saturation) if saturation > 0 else None
    h = rng.uniform(-hue, hue) if hue > 0 else None
    ops = []
    if b is not None:
        ops.append(('brightness', b))
    if s is not None:
        ops.append(('saturation', s))
    if h is not None:
        ops.append(('hue', h))
    if c is not None:
        ops.append(('contrast', c))
    rng.shuffle(ops)
    return ops


def draw_params(width, height, min_area=0.08, color=(0.4, 0.4, 0.4, 0.2), rng=random):
    """All random decisions of one VideoPrep_MSC_CJ(augment=True) call, in the reference's order (Compose runs crop, flip, jitter)."""
    crop = draw_crop(width, height, scale=(min_area, 1.), rng=rng)
    flip = draw_flip(rng=rng)
    ops = draw_jitter(*color, rng=rng)
    return dict(crop=crop, flip=flip, ops=ops)


# ---------------------------------------------------------------------------------------------------------------------------
# torchvision 0.5 functional wrappers, restated on Pillow (`tv_*`), and the clip pipeline on Pillow (`*_pil`)
# ---------------------------------------------------------------------------------------------------------------------------
def tv_resized_crop(img, i, j, h, w, size, interpolation):
    img = img.crop((j, i, j + w, i + h))                    # F.crop(img, i, j, h, w)
    return img.resize(size[::-1], interpolation)            # F.resize(img, (h, w)) -> img.resize((w, h))


def tv_adjust_brightness(img, f):
    from PIL import ImageEnhance
    return ImageEnhance.Brightness(img).enhance(f)


def tv_adjust_contrast(img, f):
    from PIL import ImageEnhance
    return ImageEnhance.Contrast(img).enhance(f)


def tv_adjust_saturation(img, f):
    from PIL import ImageEnhance
    return ImageEnhance.Color(img).enhance(f)


def hue_shift_u8(f):
    """np.uint8(hue_factor * 255) of F.adjust_hue: C conversion float -> uint8 (truncate toward zero, wrap modulo 256)."""
    return int(f * 255) & 0xFF


def tv_adjust_hue(img, f):
    from PIL import Image
    if not (-0.5 <= f <= 0.5):
        raise ValueError('hue_factor is not in [-0.5, 0.5].')
    h, s, v = img.convert('HSV').split()
    np_h = (np.array(h, dtype=np.uint8).astype(np.int32) + hue_shift_u8(f)).astype(np.uint8)      # uint8 addition wraps
    h = Image.fromarray(np_h, 'L')
    return Image.merge('HSV', (h, s, v)).convert('RGB')


_TV = dict(brightness=tv_adjust_brightness, contrast=tv_adjust_contrast, saturation=tv_adjust_saturation, hue=tv_adjust_hue)


def video_prep_pil(frames, params, crop=(224, 224), mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), normalize=True):
    """frames: (T, H, W, 3) uint8.  The reference pipeline through Pillow; returns a float32 array (3, T, crop_h, crop_w)."""
    from PIL import Image
    i, j, h, w = params['crop']
    clip = [tv_resized_crop(Image.fromarray(f), i, j, h, w, crop, Image.BILINEAR) for f in frames]
    if params['flip']:
        clip = [img.transpose(Image.FLIP_LEFT_RIGHT) for img in clip]
    out = []
    for img in clip:
        for name, f in params['ops']:
            img = _TV[name](img, f)
        out.append(np.array(img))
    return to_tensor_normalize(np.stack(out), mean, std, normalize)


def to_tensor_normalize(u8, mean, std, normalize=True):
    """ClipToTensor (float64 staging array -> .float() -> .div(255)) and Normalize (sub_ then div_, float32): (T,H,W,3) -> (3,T,H,W)."""
    x = u8.transpose(3, 0, 1, 2).astype(np.float32)
    x = x / np.float32(255.0)
    if normalize:
        m = np.asarray(mean, dtype=np.float32).reshape(3, 1, 1, 1)
        s = np.asarray(std, dtype=np.float32).reshape(3, 1, 1, 1)
        x = (x - m) / s
    return x


# ---------------------------------------------------------------------------------------------------------------------------
# numpy restatement of the Pillow arithmetic (`*_np`)
# ---------------------------------------------------------------------------------------------------------------------------
def resample_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear (triangle) filter over the whole axis (box = [0, in_size)).
    Returns (bounds (out, 2) int32 [xmin, count], kk (out, ksize) int32 fixed-point weights)."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = np.zeros(ksize, dtype=np.float64)
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            if a < 0.0:
                a = -a
            wgt = 1.0 - a if a < 1.0 else 0.0
            k[x] = wgt
            ww += wgt
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
        for x in range(ksize):
            v = k[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if k[x] < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _convolve_axis(img, bounds, kk, axis):
    """One 8bpc resampling pass along `axis` of a (..., H, W, C) uint8 array: ss = 2^21 + sum pixel * k; clip8(ss >> 22)."""
    img = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((bounds.shape[0],) + img.shape[1:], dtype=np.uint8)
    for xx in range(bounds.shape[0]):
        xmin, cnt = int(bounds[xx, 0]), int(bounds[xx, 1])
        ss = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(cnt):
            ss += img[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(ss >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resized_crop_np(frames, i, j, h, w, size):
    """crop + Image.resize(BILINEAR) for a (T, H, W, 3) uint8 clip: horizontal pass first (uint8 intermediate), then vertical
    (Resample.c ImagingResampleInner; a pass whose size does not change is skipped)."""
    x = frames[:, i:i + h, j:j + w, :]
    oh, ow = size
    if ow != w:
        bx, kx = resample_coeffs(w, ow)
        x = _convolve_axis(x, bx, kx, axis=2)
    if oh != h:
        by, ky = resample_coeffs(h, oh)
        x = _convolve_axis(x, by, ky, axis=1)
    return np.ascontiguousarray(x)


def luma_np(rgb):
    """Convert.c rgb2l: L = (R * 19595 + G * 38470 + B * 7471 + 0x8000) >> 16."""
    r, g, b = (rgb[..., k].astype(np.int64) for k in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def blend_np(deg, img, alpha):
    """Blend.c ImagingBlend(im1 = degenerate, im2 = image, float alpha) on uint8 arrays."""
    a = np.float32(alpha)
    if a == 0.0:
        return deg.copy()
    if a == 1.0:
        return img.copy()
    d, x = deg.astype(np.int32), img.astype(np.int32)
    t = d.astype(np.float32) + a * (x - d).astype(np.float32)          # float32 throughout
    if 0.0 <= a <= 1.0:
        return t.astype(np.int32).astype(np.uint8)                      # (UINT8) truncation
    return np.where(t <= 0.0, 0, np.where(t >= 255.0, 255, t.astype(np.int32))).astype(np.uint8)


def brightness_np(img, f):
    return blend_np(np.zeros_like(img), img, f)


def saturation_np(img, f):
    return blend_np(np.repeat(luma_np(img)[..., None], 3, axis=-1), img, f)


def contrast_np(img, f):
    """ImageEnhance.Contrast: degenerate = the frame's mean luma, int(mean + 0.5), per frame.  img: (..., H, W, 3) for ONE frame."""
    lum = luma_np(img)
    mean = int(lum.astype(np.int64).sum() / lum.size + 0.5)
    return blend_np(np.full_like(img, mean), img, f)


def rgb2hsv_np(rgb):
    """Convert.c rgb2hsv_row: float variables, double constants."""
    r, g, b = (rgb[..., k].astype(np.int32) for k in range(3))
    maxc, minc = np.maximum(r, np.maximum(g, b)), np.minimum(r, np.minimum(g, b))
    cr = (maxc - minc).astype(np.float32)
    safe = np.where(cr == 0, np.float32(1), cr)
    s = cr / np.where(maxc == 0, 1, maxc).astype(np.float32)
    rc = (maxc - r).astype(np.float32) / safe
    gc = (maxc - g).astype(np.float32) / safe
    bc = (maxc - b).astype(np.float32) / safe
    h = np.where(r == maxc, (bc - gc).astype(np.float32),
                 np.where(g == maxc, (2.0 + rc.astype(np.float64) - bc.astype(np.float64)).astype(np.float32),
                          (4.0 + gc.astype(np.float64) - rc.astype(np.float64)).astype(np.float32)))
    h = np.fmod(h.astype(np.float64) / 6.0 + 1.0, 1.0).astype(np.float32)
    uh = np.clip((h.astype(np.float64) * 255.0).astype(np.int32), 0, 255)
    us = np.clip((s.astype(np.float64) * 255.0).astype(np.int32), 0, 255)
    gray = maxc == minc
    return np.stack([np.where(gray, 0, uh), np.where(gray, 0, us), maxc], axis=-1).astype(np.uint8)


def hsv2rgb_np(hsv):
    """Convert.c hsv2rgb."""
    h, s, v = (hsv[..., k].astype(np.int32) for k in range(3))
    hf = h.astype(np.float32).astype(np.float64) * 6.0 / 255.0
    i = np.floor(hf).astype(np.int32)
    f = (hf - i.astype(np.float32).astype(np.float64)).astype(np.float32)
    fs = (s.astype(np.float32).astype(np.float64) / 255.0).astype(np.float32)
    vf = v.astype(np.float32).astype(np.float64)
    fs64, f64 = fs.astype(np.float64), f.astype(np.float64)
    rnd = lambda a: np.where(a >= 0, np.floor(a + 0.5), np.ceil(a - 0.5)).astype(np.int32)      # C round(): half away from zero
    p = np.clip(rnd(vf * (1.0 - fs64)), 0, 255)
    q = np.clip(rnd(vf * (1.0 - fs64 * f64)), 0, 255)
    t = np.clip(rnd(vf * (1.0 - fs64 * (1.0 - f64))), 0, 255)
    sel = i % 6
    r = np.choose(sel, [v, q, p, p, t, v])
    g = np.choose(sel, [t, v, v, q, p, p])
    b = np.choose(sel, [p, p, t, v, v, q])
    gray = s == 0
    return np.stack([np.where(gray, v, r), np.where(gray, v, g), np.where(gray, v, b)], axis=-1).astype(np.uint8)


def hue_np(img, f):
    hsv = rgb2hsv_np(img)
    hsv[..., 0] = (hsv[..., 0].astype(np.int32) + hue_shift_u8(f)).astype(np.uint8)
    return hsv2rgb_np(hsv)


def video_prep_np(frames, params, crop=(224, 224), mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), normalize=True):
    """The whole step as explicit arithmetic: what csrc/video_prep.cu computes.  frames (T, H, W, 3) uint8 -> float32 (3, T, ch, cw)."""
    i, j, h, w = params['crop']
    x = resized_crop_np(np.asarray(frames), i, j, h, w, crop)
    if params['flip']:
        x = x[:, :, ::-1, :]
    x = np.ascontiguousarray(x)
    for name, f in params['ops']:
        if name == 'brightness':
            x = brightness_np(x, f)
        elif name == 'saturation':
            x = saturation_np(x, f)
        elif name == 'hue':
            x = hue_np(x, f)
        else:
            x = np.stack([contrast_np(fr, f) for fr in x])
    return to_tensor_normalize(x, mean, std, normalize)
