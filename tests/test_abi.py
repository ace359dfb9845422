"""CPU-side checks of the C-ABI boundary: the library builds/loads and exports every symbol the header declares;
argument validation reports errors without touching a GPU; the Python layer refuses CPU tensors (no fallback)."""
import ctypes as C
import os

import pytest
import torch

from avid_cma_b200 import _lib, build


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.lib()


def test_header_symbols_are_exported_and_bound(lib):
    declared = _lib.declared_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/avid_b200.h but not exported"
    assert set(declared) == set(_lib._SIGNATURES), "ctypes signatures out of sync with the header"
    assert lib.avid_version() == 1


def test_struct_layout_matches_header(lib):
    # avid_nce_key_t: 4 x int32 + float; avid_conv_shape_t: 18 x int32
    assert C.sizeof(_lib.NceKey) == 20
    assert C.sizeof(_lib.ConvShape) == 72
    assert _lib.NceArgs.keys.offset % 4 == 0 and C.sizeof(_lib.NceArgs) % 8 == 0


def test_workspace_queries_are_host_only(lib):
    assert lib.avid_nce_workspace_bytes(64, 1024, 0, 2) > 64 * 128 * 4 * 2
    assert lib.avid_nce_workspace_bytes(0, 1024, 0, 2) == 0
    assert lib.avid_cma_topk_workspace_bytes(1000) == 1000 * 64 * 12


def test_invalid_arguments_report_einval(lib):
    rc = lib.avid_bank_update(None, None, 0, 10, None, None, None, 4, 0, 0, 0.5, 0.5, None)
    assert rc == 1 and b"NULL" in lib.avid_last_error()
    s = _lib.ConvShape(1, 1, 8, 8, 3, 1, 8, 8, 64, 1, 3, 3, 1, 1, 1, 0, 1, 1)   # ci = 3 is not a padded channel count
    rc = lib.avid_conv_forward(C.byref(s), C.c_void_p(16), C.c_void_p(16), None, C.c_void_p(16), 0, None)
    assert rc == 1 and b"ci=3" in lib.avid_last_error()
    s = _lib.ConvShape(1, 1, 8, 8, 4, 1, 8, 8, 64, 1, 3, 3, 1, 1, 1, 0, 1, 1)
    rc = lib.avid_conv_forward(C.byref(s), C.c_void_p(16), C.c_void_p(16), None, C.c_void_p(16), 7, None)
    assert rc == 4   # unknown math mode -> AVID_EUNSUPPORTED, before any launch
    # sharded optimizer: shard bounds must be float4-aligned, peers bounded, every peer buffer present (all checked before any launch)
    peers = _lib.PeerPtrs()
    peers.ptr[0] = 16
    rc = lib.avid_adam_shard_step(C.c_void_p(16), C.byref(peers), 2, C.c_void_p(16), C.c_void_p(16), 0, 8, 1, 1e-3, 0.9, 0.999, 1e-8, 0.0, 0.5, None)
    assert rc == 1 and b"rank 1 is NULL" in lib.avid_last_error()
    rc = lib.avid_adam_shard_step(C.c_void_p(16), C.byref(peers), 1, C.c_void_p(16), C.c_void_p(16), 2, 8, 1, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1.0, None)
    assert rc == 1 and b"aligned" in lib.avid_last_error()
    rc = lib.avid_pull_shards(C.c_void_p(16), C.byref(peers), 40, 0, 8, None)
    assert rc == 1 and b"world" in lib.avid_last_error()
    assert C.sizeof(_lib.PeerPtrs) == 8 * _lib.AVID_MAX_PEERS


def test_ops_refuse_cpu_tensors():
    from avid_cma_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.rows_l2_normalize_(torch.zeros(4, 128))
    from avid_cma_b200.models import R2Plus1D
    with pytest.raises(RuntimeError, match="CUDA"):
        R2Plus1D(depth=10)(torch.zeros(1, 3, 2, 16, 16))


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", os.path.join(os.path.dirname(_lib.LIB_PATH), "no_such_lib.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_model_state_dict_matches_reference_layout():
    from avid_cma_b200 import models
    from oracle import towers
    m = models.av_wrapper('R2Plus1D', {'depth': 18}, 'Conv2D', {'depth': 10}, proj_dim=[512, 512, 128])
    sd, ref = m.state_dict(), towers.state_dict_template()
    assert list(sd.keys()) == list(ref.keys())
    assert all(tuple(sd[k].shape) == tuple(ref[k].shape) for k in ref)
    assert m.out_dim == 128 and m.video_model.out_dim == 512 and m.audio_model.out_dim == 512


def test_new_entry_points_validate_arguments_on_the_host(lib):
    """Argument checks of the entry points added for CMA mining, the input side and the batched layout conversions happen before
    any launch, so they can be exercised without a GPU."""
    p16 = C.c_void_p(16)
    rc = lib.avid_log_spectrogram(p16, 2, 48000, 1000, 240, 200, 100.0, None, None, p16, p16, 64, None)      # n_fft not a power of two
    assert rc == 1 and b"power of two" in lib.avid_last_error()
    rc = lib.avid_log_spectrogram(p16, 2, 48000, 1024, 240, 500, 100.0, None, None, p16, p16, 64, None)      # more frames than the clip has
    assert rc == 1 and b"frames requested" in lib.avid_last_error()
    rc = lib.avid_log_spectrogram(p16, 2, 48000, 1024, 240, 200, 100.0, p16, None, p16, p16, 64, None)       # mean without std
    assert rc == 1
    assert lib.avid_log_spectrogram_workspace_bytes(7) == 28 and lib.avid_log_spectrogram_workspace_bytes(0) == 0
    ws = lib.avid_cma_topk_workspace_bytes(10)
    rc = lib.avid_cma_topk_certify(10, 64, 1e-3, p16, ws, None, None, None)                                    # pos_k must be < 64
    assert rc == 1 and b"certify" in lib.avid_last_error()
    rc = lib.avid_cma_topk_scan_tc(p16, p16, 10, p16, p16, 0, 100, 9, p16, ws, None)                           # unknown mode
    assert rc == 1 and b"mode" in lib.avid_last_error()
    rc = lib.avid_cma_topk_rescore(p16, p16, 10, p16, p16, 0, 100, 0, p16, ws - 1, None)                       # workspace one byte short
    assert rc == 2
    rc = lib.avid_filter_to_planes_multi(None, None, None, None, None, None, None, None, 3, None)
    assert rc == 1 and b"filter_to_planes_multi" in lib.avid_last_error()
    rc = lib.avid_cma_to_half(p16, p16, 6, None)                                                               # n must be a multiple of 4
    assert rc == 1


def test_product_code_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under avid_cma_b200/ (nor the launcher) may import it, and bench.py only in its CPU legs."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b)", re.M)
    for base, _, files in os.walk(os.path.join(root, "avid_cma_b200")):
        for f in files:
            if f.endswith(".py"):
                assert not pat.search(open(os.path.join(base, f)).read()), os.path.join(base, f)
    assert not pat.search(open(os.path.join(root, "main_avid.py")).read())
    bench = open(os.path.join(root, "bench.py")).read()
    body = bench[bench.index("def run_ours"):bench.index("if __name__")]
    calls = [m.start() for m in pat.finditer(body)]
    assert not calls, "run_ours() must not import the oracle (cpu_baseline goes through cpu_reference())"


def test_video_prep_planning_is_host_only(lib):
    """avid_video_prep_workspace_bytes validates the parameters and sizes the workspace without touching a GPU (0 = invalid)."""
    p = _lib.VideoPrep()
    p.frames, p.height, p.width = 8, 256, 340
    p.crop_top, p.crop_left, p.crop_h, p.crop_w = 10, 20, 200, 250
    p.out_h, p.out_w = 224, 224
    p.num_ops = 2
    p.op_kind[0], p.op_factor[0] = 3, 1.2       # contrast
    p.op_kind[1], p.op_factor[1] = 2, 0.1       # hue
    need = lib.avid_video_prep_workspace_bytes(C.byref(p))
    # horizontal-pass image (8 x 200 x 224 x 3) + resized image (8 x 224 x 224 x 3) + weight tables + luma sums
    assert need >= 8 * 200 * 224 * 3 + 8 * 224 * 224 * 3 and need < 2 * (8 * 200 * 224 * 3 + 8 * 224 * 224 * 3)
    arr = (_lib.VideoPrep * 3)(p, p, p)
    assert lib.avid_video_prep_batch_workspace_bytes(arr, 3) == 3 * need
    bad = _lib.VideoPrep.from_buffer_copy(p)
    bad.crop_w = 400                              # box outside the frame
    assert lib.avid_video_prep_workspace_bytes(C.byref(bad)) == 0
    assert b"crop box" in lib.avid_last_error()
    bad = _lib.VideoPrep.from_buffer_copy(p)
    bad.op_kind[1] = 3                            # two contrast ops
    assert lib.avid_video_prep_workspace_bytes(C.byref(bad)) == 0
    bad = _lib.VideoPrep.from_buffer_copy(p)
    bad.op_factor[1] = 0.75                       # hue_factor outside [-0.5, 0.5]: the reference raises too
    assert lib.avid_video_prep_workspace_bytes(C.byref(bad)) == 0 and b"hue_factor" in lib.avid_last_error()
    assert lib.avid_video_prep(None, C.byref(p), None, None, 0, None) == 1      # AVID_EINVAL


def test_conv_tc_launch_plan_is_host_only(lib):
    """avid_conv_tc_plan: the stride-parity classes of a strided input gradient partition the filter taps and the input pixels, and
    the multiply-high divisions of the tile -> pixel decode are exact on every class (no GPU involved)."""
    from avid_cma_b200 import ops
    cases = [
        # n, (t,h,w), ci, co, kernel, stride, padding
        (64, (8, 56, 56), 64, 128, (1, 3, 3), (1, 2, 2), (0, 1, 1)),      # conv3x entry of the video tower at config 2
        (64, (8, 28, 28), 128, 128, (3, 1, 1), (2, 1, 1), (1, 0, 0)),     # its temporal stride
        (64, (1, 100, 129), 64, 64, (1, 3, 3), (1, 2, 2), (0, 1, 1)),     # audio block entry: odd extents, ragged classes
        (2, (3, 7, 9), 64, 128, (3, 3, 3), (2, 2, 2), (1, 1, 1)),         # 8 classes
        (2, (4, 10, 10), 64, 128, (1, 1, 1), (2, 2, 2), (0, 0, 0)),       # 1x1x1 stride 2: 7 of 8 classes receive nothing
        (4, (8, 28, 28), 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1)),        # stride 1
    ]
    for n, (t, h, w), ci, co, k, s, p in cases:
        shape = ops.conv_shape(n, t, h, w, ci, co, k, s, p)
        out = (C.c_int64 * 5)()
        assert lib.avid_conv_tc_plan(C.byref(shape), 1, out) == 0
        ncls, mtiles, taps, pixels, div_ok = list(out)
        assert div_ok == 1
        assert taps == k[0] * k[1] * k[2]
        reached = all(kk >= ss for kk, ss in zip(k, s))
        if reached:
            assert ncls == s[0] * s[1] * s[2]
            assert pixels == n * t * h * w
        else:
            assert ncls == 1 and pixels == n * shape.to * shape.ho * shape.wo
        assert mtiles >= -(-pixels // ncls // 128)
        assert lib.avid_conv_tc_plan(C.byref(shape), 0, out) == 0
        assert list(out)[:4] == [1, -(-(n * shape.to * shape.ho * shape.wo) // 128), k[0] * k[1] * k[2], n * shape.to * shape.ho * shape.wo] and out[4] == 1
    assert lib.avid_conv_tc_plan(None, 1, None) != 0
