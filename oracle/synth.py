"""Deterministic synthetic inputs for the parity configs.

Everything is drawn from numpy's MT19937 `RandomState`, whose streams are stable
across numpy / torch versions, so the golden generator (which runs the imported
reference in the build container), the oracle and the CUDA tests all see bit-identical
inputs without shipping multi-megabyte fixtures.
"""
import zlib
import numpy as np
import torch


def _rs(seed, tag=""):
    return np.random.RandomState((int(seed) * 1000003 + zlib.crc32(tag.encode())) % (2 ** 32))


def normal(shape, seed, tag):
    return torch.from_numpy(_rs(seed, tag).standard_normal(size=shape).astype(np.float32))


def clips(batch, frames=8, size=112, seed=0):
    """video (B,3,T,H,W) fp32, as datasets/video_db.py:219-265 would hand to the model."""
    return normal((batch, 3, frames, size, size), seed, "video")


def spectrograms(batch, t=100, f=129, seed=0):
    """audio (B,1,T,F) fp32 log-spectrograms."""
    return normal((batch, 1, t, f), seed, "audio")


def bank(num_rows, dim=128, seed=0, tag="bank_v"):
    """Row-normalised N(0,1) bank, the distribution of init_memory (avid.py:88-96)."""
    x = _rs(seed, tag).standard_normal(size=(num_rows, dim)).astype(np.float32)
    x /= np.maximum(np.sqrt((x.astype(np.float64) ** 2).sum(1, keepdims=True)), 1e-12).astype(np.float32)
    return torch.from_numpy(x)


def instance_ids(batch, num_rows, seed=0):
    """Distinct instance indices y (B,) int64."""
    return torch.from_numpy(_rs(seed, "y").permutation(num_rows)[:batch].astype(np.int64))


def negatives(y, num_neg, num_rows, seed=0):
    """(B,K) int64 uniform over [0,N) minus {y_b}: avid.py:82-86 with an injected draw."""
    r = _rs(seed, "neg").randint(0, num_rows - 1, size=(y.shape[0], num_neg)).astype(np.int64)
    r = torch.from_numpy(r)
    return r + (r >= y.view(-1, 1)).long()


def raw_negatives(batch, num_neg, upper, seed=0):
    """(B,K) int64 uniform over [0, upper): the raw alias-method draw before any remap."""
    return torch.from_numpy(_rs(seed, "rawneg").randint(0, upper, size=(batch, num_neg)).astype(np.int64))


def embeddings(batch, dim=128, seed=0):
    return normal((batch, dim), seed, "emb_v"), normal((batch, dim), seed, "emb_a")


def fill_state_dict(state_dict, seed=0):
    """Deterministic weights for a model state_dict with the reference's 267 keys.

    conv / linear weights ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (the scale of PyTorch's
    default init), BN affine perturbed away from (1, 0) so the affine path is exercised,
    running stats at their defaults.  Returns a new dict of fp32 tensors.
    """
    out = {}
    for k in state_dict:
        v = state_dict[k]
        shape = tuple(v.shape)
        rs = _rs(seed, k)
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros(shape, dtype=torch.long)
        elif k.endswith("running_mean"):
            out[k] = torch.zeros(shape)
        elif k.endswith("running_var"):
            out[k] = torch.ones(shape)
        elif len(shape) >= 2:  # conv / linear weight
            fan_in = int(np.prod(shape[1:]))
            b = 1.0 / np.sqrt(fan_in)
            out[k] = torch.from_numpy(rs.uniform(-b, b, size=shape).astype(np.float32))
        elif k.endswith("weight"):  # BN gamma
            out[k] = torch.from_numpy((1.0 + 0.1 * rs.uniform(-1, 1, size=shape)).astype(np.float32))
        elif ".projection." in k:  # linear bias
            out[k] = torch.from_numpy(rs.uniform(-0.04, 0.04, size=shape).astype(np.float32))
        else:  # BN beta
            out[k] = torch.from_numpy((0.1 * rs.uniform(-1, 1, size=shape)).astype(np.float32))
    return out
