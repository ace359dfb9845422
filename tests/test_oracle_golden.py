"""Pin the oracle (oracle/*.py) against vectors produced by the imported reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import criterion as oc
from oracle import synth, towers


def _close(a, b, rtol, atol=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


@pytest.mark.parametrize("tag", ["cross", "joint", "cfg1"])
def test_avid_criterion_matches_reference(golden, tag):
    g = golden("criterion_" + tag)
    N, B, K, seed = int(g["N"]), int(g["B"]), int(g["K"]), int(g["seed"])
    mom = g["momentum"].tolist()
    keys = oc.avid_keys(K, float(g["xModal"]), float(g["wModal"]))
    bv, ba = synth.bank(N, seed=seed, tag="bank_v"), synth.bank(N, seed=seed, tag="bank_a")
    Z = -1.0
    for s in range(int(g["steps"])):
        ev, ea = synth.embeddings(B, seed=seed + 100 * s)
        y = synth.instance_ids(B, N, seed=seed + 100 * s)
        idx = synth.negatives(y, K, N, seed=seed + 100 * s)
        r = oc.criterion_forward_backward(ev, ea, y, bv, ba, idx, keys, Z)
        Z = r["Z"]
        _close(Z, g[f"s{s}_Z"], 1e-6)
        _close(r["total"], g[f"s{s}_total"], 2e-6)
        for k in keys:
            _close(r["losses"][k.name], g[f"s{s}_Loss/{k.name}"], 2e-6)
        _close(r["grad_v"], g[f"s{s}_grad_v"], 1e-4, 1e-7)
        _close(r["grad_a"], g[f"s{s}_grad_a"], 1e-4, 1e-7)
        oc.bank_update(bv, ba, ev, ea, y, mom)
        _close(bv[y], g[f"s{s}_rows_v"], 1e-5, 1e-7)
        _close(ba[y], g[f"s{s}_rows_a"], 1e-5, 1e-7)


@pytest.mark.parametrize("tag,mode", [("consensus", "consensus"), ("union", "union")])
def test_cma_matches_reference(golden, tag, mode):
    g = golden("cma_" + tag)
    N, B, K, pos_k, seed = int(g["N"]), int(g["B"]), int(g["K"]), int(g["pos_k"]), int(g["seed"])
    Kw = None if int(g["Kw"]) < 0 else int(g["Kw"])
    bv, ba = synth.bank(N, seed=seed, tag="bank_v"), synth.bank(N, seed=seed, tag="bank_a")
    pos = oc.cma_topk(bv, ba, pos_k, mode)
    assert np.array_equal(pos.numpy(), g["positive_set"])
    ev, ea = synth.embeddings(B, seed=seed)
    y = synth.instance_ids(B, N, seed=seed)
    raw = synth.raw_negatives(B, K, N - pos_k, seed=seed)
    neg = oc.remap_negatives_cma(raw, pos[y])
    assert np.array_equal(neg.numpy(), g["neg_idx"])
    keys = oc.avid_cma_keys(K, Kw)
    r = oc.criterion_forward_backward(ev, ea, y, bv, ba, neg, keys, -1.0, positive_set=pos)
    _close(r["Z"], g["Z"], 1e-6)
    _close(r["total"], g["total"], 2e-6)
    for k in keys:
        _close(r["losses"][k.name], g["Loss/" + k.name], 2e-6)
    _close(r["grad_v"], g["grad_v"], 1e-4, 1e-7)
    _close(r["grad_a"], g["grad_a"], 1e-4, 1e-7)
    oc.bank_update(bv, ba, ev, ea, y, 0.5)
    _close(bv[y], g["rows_v"], 1e-5, 1e-7)


def test_state_dict_template_has_reference_keys(golden):
    g = golden("step_config1")
    sd = towers.state_dict_template()
    assert len(sd) == 267
    assert sorted(towers.param_keys(sd)) == sorted(g["grad_names"].tolist())


def test_full_step_matches_reference(golden):
    """BASELINE config 1 through the oracle towers + criterion, fp32, vs the imported reference."""
    g = golden("step_config1")
    B, N, K, size, seed = int(g["B"]), int(g["N"]), int(g["K"]), int(g["size"]), int(g["seed"])
    spec = g["spec"].tolist()
    torch.set_num_threads(8)
    sd = synth.fill_state_dict(towers.state_dict_template(), seed=seed)
    params = {k: sd[k].requires_grad_(True) for k in towers.param_keys(sd)}
    video, audio = synth.clips(B, 8, size, seed), synth.spectrograms(B, spec[0], spec[1], seed)
    y = torch.from_numpy(g["y"])
    idx = synth.negatives(y, K, N, seed)
    bv, ba = synth.bank(N, seed=seed, tag="bank_v"), synth.bank(N, seed=seed, tag="bank_a")
    ve, ae = towers.av_forward(video, audio, sd, training=True)
    total, losses, Z = oc.criterion_forward(ve, ae, y, bv, ba, idx, oc.avid_keys(K))
    total.backward()
    _close(ve.detach(), g["video_emb"], 1e-4, 1e-6)
    _close(ae.detach(), g["audio_emb"], 1e-4, 1e-6)
    _close(total.detach(), g["total"], 1e-5)
    _close(Z, g["Z"], 1e-5)
    norms = dict(zip(g["grad_names"].tolist(), g["grad_norms"].tolist()))
    for k, p in params.items():
        assert abs(float(p.grad.double().norm()) - norms[k]) <= 2e-3 * norms[k] + 1e-9, k
    for k in g.files:
        if k.startswith("grad::"):
            _close(params[k[6:]].grad, g[k], 2e-3, 2e-6 * np.abs(g[k]).max())
        elif k.startswith("rm::"):
            _close(sd[k[4:] + ".running_mean"], g[k], 1e-4, 1e-7)
        elif k.startswith("rv::"):
            _close(sd[k[4:] + ".running_var"], g[k], 1e-4, 1e-7)
    oc.bank_update(bv, ba, ve, ae, y, 0.5)
    _close(bv[y], g["rows_v"], 1e-4, 1e-6)
    _close(ba[y], g["rows_a"], 1e-4, 1e-6)


def test_audio_oracle_stft_matches_scipy():
    """oracle/audio.py restates librosa.stft (absent here); cross-check it against scipy's independent STFT with the same
    conventions (periodic hann window, reflect-padded centred frames) and check the bin pairing / dB floor arithmetic."""
    import numpy as np
    import scipy.signal
    from oracle import audio as oa
    sig = np.random.default_rng(0).standard_normal(5000)
    P = oa.stft_power(sig, 1024, 240)
    _, _, Z = scipy.signal.stft(np.pad(sig, 512, mode="reflect"), window="hann", nperseg=1024, noverlap=1024 - 240, boundary=None,
                                padded=False, return_onesided=True)
    Z = Z * scipy.signal.get_window("hann", 1024, fftbins=True).sum()
    assert P.shape == (513, 1 + 5000 // 240) and np.abs(np.abs(Z) ** 2 - P).max() < 1e-12 * P.max()
    out = oa.log_spectrogram(sig, 24000, n_fft=512, hop_size=0.01, duration=0.1)
    assert out.shape == (1, 10, 257) and out.max() - out.min() <= 100.0
    pair = 0.5 * (P[1:, :10].reshape(256, 2, -1)).sum(1)
    np.testing.assert_allclose(out[0, :, 1:], 10 * np.log10(np.maximum(pair, 1e-10)).T, rtol=1e-12)


@pytest.mark.parametrize("mode", ["video", "audio"])
def test_cma_single_modality_mining_matches_reference(golden, mode):
    """sampling type 'video' / 'audio' (avid_cma.py:60-63): positives ranked by one modality's similarity only."""
    g = golden("cma_mining_" + mode)
    N, pos_k, seed = int(g["N"]), int(g["pos_k"]), int(g["seed"])
    bv, ba = synth.bank(N, seed=seed, tag="bank_v"), synth.bank(N, seed=seed, tag="bank_a")
    assert np.array_equal(oc.cma_topk(bv, ba, pos_k, mode).numpy(), g["positive_set"])
