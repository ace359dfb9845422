"""Tensor-level wrappers over the C ABI: validate torch CUDA tensors, pass raw pointers + the current stream.

PyTorch is plumbing here (device memory and streams); every function below launches only kernels of
libavid_b200.so.  All tensors must be contiguous CUDA tensors on the current device.
"""
import ctypes as C

import os

import torch

from . import _lib
from ._lib import ConvShape, NceArgs, check

MATH_FP32, MATH_BF16X3, MATH_BF16 = _lib.MATH_FP32, _lib.MATH_BF16X3, _lib.MATH_BF16
BN_EPS, BN_MOMENTUM = 1e-5, 0.1


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t, dtype=torch.float32, optional=False):
    if t is None:
        if optional:
            return None
        raise ValueError("tensor is None")
    if not t.is_cuda:
        raise RuntimeError("avid_cma_b200 ops need CUDA tensors (there is no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return C.c_void_p(t.data_ptr())


# ---- optional per-launch timing (bench.py roofline): CUDA events on the launching stream around each hot kernel
_prof = None


_prof_all = False


def profile_begin(all_families=None):
    """Start recording CUDA-event pairs around the tensor-core conv launches and the fused criterion (bench.py roofline).
    all_families (or AVID_PROFILE_ALL=1): also time the BatchNorm / layout / optimizer wrappers -- ~4x more events per step, for
    diagnostic runs only (the event records themselves cost about a millisecond per step)."""
    global _prof, _prof_all
    _prof = []
    _prof_all = os.environ.get("AVID_PROFILE_ALL", "0") == "1" if all_families is None else bool(all_families)


def profile_end():
    """[(family, algorithmic work (FLOP or bytes), milliseconds)] for every instrumented launch since profile_begin()."""
    global _prof
    rec, _prof = _prof or [], None
    torch.cuda.synchronize()
    return [(name, work, e0.elapsed_time(e1)) for name, work, e0, e1 in rec]


def _t0():
    if _prof is None:
        return None
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _t1(e0, name, work):
    if e0 is not None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        _prof.append((name, work, e0, e1))


def _timed(name):
    """Time a wrapper's launches as family `name` when bench.py profiles a step (no algorithmic work attached)."""
    def deco(fn):
        def inner(*a, **k):
            if not _prof_all:
                return fn(*a, **k)
            e0 = _t0()
            r = fn(*a, **k)
            _t1(e0, name, 0.0)
            return r
        inner.__name__, inner.__doc__ = fn.__name__, fn.__doc__
        return inner
    return deco


# ---- filter gradients on a side stream --------------------------------------------------------------------------------------
# The filter gradient of a layer is off the critical path of the backward pass (nothing but the optimizer reads it), while the
# BatchNorm-backward apply of the NEXT layer down is on it and is HBM-bound with no shared memory.  With AVID_WGRAD_STREAM=1 (default) every
# tensor-core filter gradient is enqueued on a side stream behind an event recorded AFTER the input gradient of its layer: the persistent
# wgrad CTAs start once the input gradient has drained, together with the BatchNorm pass that follows it on the main stream
# (measured A/B in one run: 24.71 / 24.97 ms -> 24.31 / 24.65 ms per step).
_side_streams = {}
_side_keep = []


def wgrad_stream_enabled():
    return os.environ.get("AVID_WGRAD_STREAM", "1") == "1" and _prof is None


def side_event():
    """An event on the current stream: the point a side launch has to wait for."""
    ev = torch.cuda.Event()
    ev.record()
    return ev


def side_run(ev, fn, keep=()):
    """Run fn() on the side stream of the current stream once `ev` has happened; `keep`: tensors fn reads (held until side_join, so the
    caching allocator cannot hand their memory to later launches of the main stream)."""
    cur = torch.cuda.current_stream()
    side = _side_streams.get(cur.cuda_stream)
    if side is None:
        side = _side_streams[cur.cuda_stream] = torch.cuda.Stream()
    side.wait_event(ev)
    with torch.cuda.stream(side):
        out = fn()
    _side_keep.append(keep)
    return out


def side_join():
    """The current stream waits for everything enqueued on its side stream."""
    cur = torch.cuda.current_stream()
    side = _side_streams.get(cur.cuda_stream)
    if side is not None:
        cur.wait_stream(side)
    _side_keep.clear()


class ZeroArena:
    """Zero-initialised scratch for one forward or backward pass of a tower: the accumulators the kernels add into with
    atomics (filter gradients, BatchNorm sums) come from ONE buffer zeroed by ONE memset instead of a fill kernel each.
    The first pass measures the demand (and falls back to torch.zeros); later passes reuse the buffer."""

    def __init__(self):
        self.buf, self.capacity, self.pos, self.demand = None, 0, 0, 0

    def begin(self, device):
        if self.demand > self.capacity:
            self.buf = torch.empty(self.demand, dtype=torch.uint8, device=device)
            self.capacity = self.demand
        if self.buf is not None:
            check(_lib.lib().avid_zero_bytes(C.c_void_p(self.buf.data_ptr()), self.capacity, _stream()))
        self.pos, self.demand = 0, 0

    def zeros(self, shape, dtype, device):
        n = 1
        for d in shape:
            n *= d
        nbytes = n * torch.empty(0, dtype=dtype).element_size()
        aligned = (nbytes + 255) & ~255
        self.demand += aligned
        if self.buf is not None and self.pos + aligned <= self.capacity and self.buf.device == device:
            out = self.buf[self.pos:self.pos + nbytes].view(dtype).view(shape)
            self.pos += aligned
            return out
        return torch.zeros(shape, dtype=dtype, device=device)


_arena = None


def set_arena(arena):
    """Route the accumulator allocations of the following kernel wrappers through `arena` (None: plain torch.zeros)."""
    global _arena
    _arena = arena


def _zeros(shape, dtype, device):
    if _arena is not None:
        return _arena.zeros(tuple(shape), dtype, torch.device(device))
    return torch.zeros(shape, dtype=dtype, device=device)


def _conv_flops(s, ci_real=None):
    return 2.0 * s.n * s.to * s.ho * s.wo * s.kt * s.kh * s.kw * (ci_real or s.ci) * s.co


_replayed_launches = 0


def launch_count():
    """Kernels of libavid_b200.so launched so far: direct launches (counted by the library) + launches replayed from CUDA graphs."""
    return int(_lib.lib().avid_launch_count()) + _replayed_launches


def add_launches(n):
    """A CUDA-graph replay executes `n` captured launches the library's own counter does not see."""
    global _replayed_launches
    _replayed_launches += int(n)


def reset_launch_count():
    global _replayed_launches
    _replayed_launches = 0
    _lib.lib().avid_reset_launch_count()


def profiling():
    """True while bench.py records CUDA-event pairs around the launches (event timing cannot be captured into a graph)."""
    return _prof is not None


# ------------------------------------------------------------------ criterion

def nce_workspace(batch, num_neg, pos_k, num_keys, device):
    n = int(_lib.lib().avid_nce_workspace_bytes(batch, num_neg, max(pos_k, 0), num_keys))
    return torch.zeros(n, dtype=torch.uint8, device=device)     # the ticket counters at its start must be zero before the first call


def _group_layout(name, t, elem_bytes, rows_per_group):
    """(pointer to record 0, bytes between records) of a packed per-rank tensor (W, rows_per_group, ...) whose records are dense."""
    if t.shape[1] != rows_per_group or not t[0].is_contiguous():
        raise ValueError(f"{name}: every rank record must be a dense block of {rows_per_group} rows")
    # a single record has no stride to speak of (PyTorch reports an arbitrary one for a size-1 dimension)
    return t[0], (t.stride(0) * elem_bytes if t.shape[0] > 1 else None)


def make_nce_args(emb_v, emb_a, y, bank_v, bank_a, keys, num_neg, Z, *, num_rows=None, row_begin=0, row_end=None,
                  neg_idx=None, seed=0, offset=0, positive_set=None, mean_batch=0, temperature=0.07,
                  loss_keys=None, loss_total=None, grad_v=None, grad_a=None, scores=None, neg_idx_out=None,
                  grad_hat_v=None, grad_hat_a=None, loss_part=None, bad_index=None):
    """keys: sequence of (ctx, bank, pos_mode, num_neg, weight).

    Dense call: emb_* (B,128), y (B), neg_idx (B,K), grad_hat_* (B,128), loss_part (num_keys,B).
    Packed call (sharded protocol): emb_* (W,B,128), y (W,B), neg_idx (W,B,K) are strided views into the ONE all-gathered
    buffer of the step (a dense record per rank), grad_hat_* (W,B,128) and loss_part (W,num_keys,B) views into the ONE buffer
    that is reduce-scattered afterwards: the kernel addresses rank records by stride, nothing is re-packed."""
    a = NceArgs()
    grouped = emb_v.dim() == 3
    views = {}
    if grouped:
        W, B = emb_v.shape[0], emb_v.shape[1]
        ins = [("emb_v", emb_v, 4), ("emb_a", emb_a, 4), ("y", y, 8)] + ([("neg_idx", neg_idx, 8)] if neg_idx is not None else [])
        strides = set()
        for name, t, eb in ins:
            views[name], st = _group_layout(name, t, eb, B)
            strides.add(st)
        strides.discard(None)
        if len(strides) > 1:
            raise ValueError("packed inputs must share one record stride")
        a.group_batch, a.in_group_stride = B, (strides.pop() if strides else 0)
        outs = [(n_, t) for n_, t in (("grad_hat_v", grad_hat_v), ("grad_hat_a", grad_hat_a)) if t is not None]
        strides = set()
        for name, t in outs:
            views[name], st = _group_layout(name, t, 4, B)
            strides.add(st)
        if loss_part is not None:
            if tuple(loss_part.shape) != (W, len(keys), B) or not loss_part[0].is_contiguous():
                raise ValueError("packed loss_part must be (W, num_keys, B) with dense records")
            views["loss_part"] = loss_part[0]
            strides.add(loss_part.stride(0) * 4 if W > 1 else None)
        strides.discard(None)
        if len(strides) > 1:
            raise ValueError("packed outputs must share one record stride")
        a.out_group_stride = strides.pop() if strides else 0
        batch = W * B
        e_v, e_a, y_, n_ = views["emb_v"], views["emb_a"], views["y"], views.get("neg_idx")
        gh_v, gh_a, lp = views.get("grad_hat_v"), views.get("grad_hat_a"), views.get("loss_part")
    else:
        batch = emb_v.shape[0]
        e_v, e_a, y_, n_, gh_v, gh_a, lp = emb_v, emb_a, y, neg_idx, grad_hat_v, grad_hat_a, loss_part
        if emb_v.shape != emb_a.shape or emb_v.shape[1] != 128 or y.shape[0] != emb_v.shape[0]:
            raise ValueError("embeddings must be (B,128) and y (B,)")
        if neg_idx is not None and tuple(neg_idx.shape) != (batch, num_neg):
            raise ValueError("neg_idx must be (B,K)")
    a.emb_video, a.emb_audio, a.y = _p(e_v), _p(e_a), _p(y_, torch.int64)
    a.bank_video, a.bank_audio = _p(bank_v), _p(bank_a)
    a.num_rows = bank_v.shape[0] if num_rows is None else num_rows
    a.row_begin = row_begin
    a.row_end = a.num_rows if row_end is None else row_end
    if bank_v.shape[0] != a.row_end - a.row_begin or bank_a.shape != bank_v.shape or bank_v.shape[1] != 128:
        raise ValueError("bank shapes do not match the row range / embedding width 128")
    a.batch, a.mean_batch, a.num_neg = batch, mean_batch, num_neg
    a.neg_idx = _p(n_, torch.int64, optional=True)
    a.seed, a.offset = seed, offset
    a.positive_set = _p(positive_set, torch.int32, optional=True)
    a.pos_k = positive_set.shape[1] if positive_set is not None else 0
    a.num_keys = len(keys)
    for i, (ctx, bank, pos_mode, kn, w) in enumerate(keys):
        a.keys[i].ctx, a.keys[i].bank, a.keys[i].pos_mode, a.keys[i].num_neg, a.keys[i].weight = ctx, bank, pos_mode, kn, w
    a.avg_exp_score = _p(Z, optional=True)
    a.temperature = temperature
    a.loss_keys, a.loss_total = _p(loss_keys, optional=True), _p(loss_total, optional=True)
    a.grad_video, a.grad_audio = _p(grad_v, optional=True), _p(grad_a, optional=True)
    a.scores = _p(scores, optional=True)
    a.neg_idx_out = _p(neg_idx_out, torch.int64, optional=True)
    a.grad_hat_video, a.grad_hat_audio = _p(gh_v, optional=True), _p(gh_a, optional=True)
    a.loss_part = _p(lp, optional=True)
    # a pinned host int32 the kernel raises when an instance index is out of range (polled by the criterion without a sync)
    a.bad_index = C.c_void_p(bad_index.data_ptr()) if bad_index is not None else None
    # the struct only carries raw pointers: keep the tensors alive as long as the args object
    a.tensors = dict(emb_v=emb_v, emb_a=emb_a, y=y, bank_v=bank_v, bank_a=bank_a, neg_idx=neg_idx, positive_set=positive_set, Z=Z,
                     loss_keys=loss_keys, loss_total=loss_total, grad_v=grad_v, grad_a=grad_a, scores=scores, neg_idx_out=neg_idx_out,
                     grad_hat_v=grad_hat_v, grad_hat_a=grad_hat_a, loss_part=loss_part, bad_index=bad_index)
    return a


def nce_forward_backward(args, workspace):
    e0 = _t0()
    nb = len({args.keys[i].bank for i in range(args.num_keys)})
    check(_lib.lib().avid_nce_forward_backward(C.byref(args), _p(workspace, torch.uint8), workspace.numel(), _stream()))
    # bytes: every scored row is read once; a shard holds (row_end - row_begin) / num_rows of the rows of every query
    held = (args.row_end - args.row_begin) / float(args.num_rows)
    _t1(e0, "nce_fused", float(nb * args.batch * (args.num_neg + 1 + args.pos_k) * 512) * held)


def nce_finalize(args, workspace):
    check(_lib.lib().avid_nce_finalize(C.byref(args), _p(workspace, torch.uint8), workspace.numel(), _stream()))


def nce_partition_mean(args, key, out, workspace):
    check(_lib.lib().avid_nce_partition_mean(C.byref(args), key, _p(out), _p(workspace, torch.uint8), workspace.numel(), _stream()))


def bank_update(bank_v, bank_a, emb_v, emb_a, y, mom_v, mom_a, row_begin=0, row_end=None):
    """emb_* (n,128) and y (n) dense, or the packed views (W,B,128) / (W,B) of the step's all-gathered buffer."""
    row_end = row_begin + bank_v.shape[0] if row_end is None else row_end
    n, gb, stride = y.shape[0], 0, 0
    if emb_v.dim() == 3:
        n, gb = emb_v.shape[0] * emb_v.shape[1], emb_v.shape[1]
        (emb_v, s0), (emb_a, s1), (y, s2) = (_group_layout("emb_v", emb_v, 4, gb), _group_layout("emb_a", emb_a, 4, gb),
                                             _group_layout("y", y, 8, gb))
        if len({s0, s1, s2} - {None}) > 1:
            raise ValueError("packed inputs must share one record stride")
        stride = s0 or 0
    check(_lib.lib().avid_bank_update(_p(bank_v), _p(bank_a), row_begin, row_end, _p(emb_v), _p(emb_a), _p(y, torch.int64),
                                      n, gb, stride, float(mom_v), float(mom_a), _stream()))


def bank_init_(bank, row_begin, seed, which):
    """init_memory (avid.py:88-96) for the rows [row_begin, row_begin + len(bank)) of bank `which`: L2-normalised N(0,1) rows that
    depend only on (seed, which, row) -- no broadcast from rank 0 needed."""
    check(_lib.lib().avid_bank_init(_p(bank), row_begin, bank.shape[0], seed & 0xFFFFFFFFFFFFFFFF, which, _stream()))
    return bank


def rows_l2_normalize_(x):
    check(_lib.lib().avid_rows_l2_normalize(_p(x), x.shape[0], _stream()))
    return x


def sample_negatives(y, num_neg, num_rows, seed, offset, positive_set=None):
    out = torch.empty(y.shape[0], num_neg, dtype=torch.int64, device=y.device)
    check(_lib.lib().avid_sample_negatives(_p(y, torch.int64), y.shape[0], num_neg, num_rows, _p(positive_set, torch.int32, optional=True),
                                           positive_set.shape[1] if positive_set is not None else 0, seed, offset, _p(out, torch.int64), _stream()))
    return out


CMA_MODES = {"consensus": 0, "union": 1, "video": 2, "audio": 3}


CMA_EPS = 1.0e-3      # rigorous bound of |fp16-input similarity - exact similarity| for unit rows (2^-10 + fp32 accumulation slack)


def _cma_to_half(x):
    out = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    check(_lib.lib().avid_cma_to_half(_p(x), _p(out, torch.float16), x.numel(), _stream()))
    return out


def _shard_iter(cand_shards):
    return cand_shards() if callable(cand_shards) else cand_shards


class _CmaExactEngine:
    """fp32 CUDA-core search (csrc/cma.cu): begin / one scan per candidate shard / finish."""

    def __init__(self, q_video, q_audio, pos_k, mode):
        self.qv, self.qa, self.pos_k, self.mode = q_video, q_audio, pos_k, CMA_MODES[mode]
        self.nq = q_video.shape[0]
        L = _lib.lib()
        self.ws = torch.empty(int(L.avid_cma_topk_workspace_bytes(self.nq)), dtype=torch.uint8, device=q_video.device)
        check(L.avid_cma_topk_begin(self.nq, _p(self.ws, torch.uint8), self.ws.numel(), _stream()))

    def scan(self, cv, ca, begin):
        check(_lib.lib().avid_cma_topk_scan(_p(self.qv), _p(self.qa), self.nq, _p(cv), _p(ca), begin, cv.shape[0], self.mode, self.pos_k,
                                            _p(self.ws, torch.uint8), self.ws.numel(), _stream()))

    def finish(self):
        out = torch.empty(self.nq, self.pos_k, dtype=torch.int32, device=self.qv.device)
        check(_lib.lib().avid_cma_topk_finish(self.nq, self.pos_k, _p(out, torch.int32), _p(self.ws, torch.uint8), self.ws.numel(), _stream()))
        return out


class _CmaTensorCoreEngine(_CmaExactEngine):
    """tcgen05 candidate generation + exact re-scoring per shard, certificate at the end (csrc/cma_tc.cu).
    finish() -> (positives, rows of the queries without a certificate)."""

    def __init__(self, q_video, q_audio, pos_k, mode, eps):
        super().__init__(q_video, q_audio, pos_k, mode)
        self.eps = float(eps)
        self.qv_h, self.qa_h = _cma_to_half(q_video), _cma_to_half(q_audio)

    def scan(self, cv, ca, begin):
        L = _lib.lib()
        same = cv.data_ptr() == self.qv.data_ptr() and cv.shape[0] == self.nq
        cv_h, ca_h = (self.qv_h, self.qa_h) if same else (_cma_to_half(cv), _cma_to_half(ca))
        wp, wn = _p(self.ws, torch.uint8), self.ws.numel()
        check(L.avid_cma_topk_scan_tc(_p(self.qv_h, torch.float16), _p(self.qa_h, torch.float16), self.nq, _p(cv_h, torch.float16),
                                      _p(ca_h, torch.float16), begin, cv.shape[0], self.mode, wp, wn, _stream()))
        check(L.avid_cma_topk_rescore(_p(self.qv), _p(self.qa), self.nq, _p(cv), _p(ca), begin, cv.shape[0], self.mode, wp, wn, _stream()))

    def finish(self):
        dev = self.qv.device
        fail_count = torch.zeros(1, dtype=torch.int32, device=dev)
        fail_list = torch.empty(self.nq, dtype=torch.int32, device=dev)
        check(_lib.lib().avid_cma_topk_certify(self.nq, self.pos_k, self.eps, _p(self.ws, torch.uint8), self.ws.numel(),
                                               _p(fail_count, torch.int32), _p(fail_list, torch.int32), _stream()))
        out = super().finish()
        return out, fail_list[:int(fail_count.item())].long()


def _cma_topk_exact(q_video, q_audio, cand_shards, pos_k, mode, run=True):
    """One walk over the shards with the fp32 engine.  run=False only walks them (a rank without queries to re-mine must still
    take part in the collectives that stream the shards)."""
    eng = _CmaExactEngine(q_video, q_audio, pos_k, mode) if run else None
    for cv, ca, begin in _shard_iter(cand_shards):
        if run:
            eng.scan(cv, ca, begin)
    return eng.finish() if run else None


def cma_topk(q_video, q_audio, cand_shards, pos_k, mode="consensus", exact=None, any_rank=None, eps=CMA_EPS, stats=None):
    """CMASampler.sample_instance for all query rows (avid_cma.py:42-73).  cand_shards: a list of (cand_video, cand_audio,
    cand_begin), or a callable returning a fresh iterator of them (sharded runs stream the shards through collectives).
    Returns (Q, pos_k) int32 sorted positives.

    Default: tensor-core candidate generation + exact re-scoring + certificate (csrc/cma_tc.cu); queries without a certificate
    are re-mined with the fp32 kernel, which needs a second walk over the shards -- `any_rank(flag) -> flag` must OR the flag
    over the ranks when the shard iterator contains collectives.  exact=True (or AVID_CMA_EXACT=1): fp32 kernel only."""
    if mode not in CMA_MODES:
        raise ValueError(mode)
    if exact is None:
        exact = os.environ.get("AVID_CMA_EXACT", "0") == "1"
    if exact:
        return _cma_topk_exact(q_video, q_audio, cand_shards, pos_k, mode)
    eng = _CmaTensorCoreEngine(q_video, q_audio, pos_k, mode, eps)
    for cv, ca, begin in _shard_iter(cand_shards):
        eng.scan(cv, ca, begin)
    out, rows = eng.finish()
    n_fail = int(rows.numel())
    if stats is not None:
        stats["uncertified"] = n_fail
    again = n_fail > 0
    if any_rank is not None:
        again = bool(any_rank(again))
    if again:
        if not (callable(cand_shards) or isinstance(cand_shards, (list, tuple))):
            raise RuntimeError("cma_topk: re-mining needs a re-iterable shard source (pass a list or a callable)")
        if n_fail > 0:
            out[rows] = _cma_topk_exact(q_video[rows].contiguous(), q_audio[rows].contiguous(), cand_shards, pos_k, mode)
        else:
            _cma_topk_exact(None, None, cand_shards, pos_k, mode, run=False)
    return out


# ------------------------------------------------------------------ encoders

def conv_shape(n, ti, hi, wi, ci, co, kernel, stride, padding):
    kt, kh, kw = kernel
    st, sh, sw = stride
    pt, ph, pw = padding
    to, ho, wo = (ti + 2 * pt - kt) // st + 1, (hi + 2 * ph - kh) // sh + 1, (wi + 2 * pw - kw) // sw + 1
    return ConvShape(n, ti, hi, wi, ci, to, ho, wo, co, kt, kh, kw, st, sh, sw, pt, ph, pw)


def conv_forward(s, x, w_tap, addend=None, out=None, math=MATH_FP32, ci_real=None):
    if out is None:
        out = torch.empty(s.n, s.to, s.ho, s.wo, s.co, dtype=torch.float32, device=x.device)
    assert x.numel() == s.n * s.ti * s.hi * s.wi * s.ci and w_tap.numel() >= s.kt * s.kh * s.kw * s.ci * s.co
    e0 = _t0()
    check(_lib.lib().avid_conv_forward(C.byref(s), _p(x), _p(w_tap), _p(addend, optional=True), _p(out), math, _stream()))
    _t1(e0, "conv_forward", _conv_flops(s, ci_real))
    return out


def conv_dgrad(s, dout, w_tap_t, addend=None, out=None, math=MATH_FP32):
    if out is None:
        out = torch.empty(s.n, s.ti, s.hi, s.wi, s.ci, dtype=torch.float32, device=dout.device)
    assert dout.numel() == s.n * s.to * s.ho * s.wo * s.co
    e0 = _t0()
    check(_lib.lib().avid_conv_dgrad(C.byref(s), _p(dout), _p(w_tap_t), _p(addend, optional=True), _p(out), math, _stream()))
    _t1(e0, "conv_dgrad", _conv_flops(s))
    return out


def conv_wgrad(s, x, dout, math=MATH_FP32, ci_real=None):
    dw = _zeros((s.kt * s.kh * s.kw, s.ci, s.co), torch.float32, x.device)
    e0 = _t0()
    check(_lib.lib().avid_conv_wgrad(C.byref(s), _p(x), _p(dout), _p(dw), math, _stream()))
    _t1(e0, "conv_wgrad", _conv_flops(s, ci_real))
    return dw


def split_bf16(x, need_lo=True):
    """fp32 tensor -> (hi, lo) bf16 planes with x ~= hi + lo (lo is None when need_lo is False)."""
    hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if need_lo else None
    check(_lib.lib().avid_split_bf16(_p(x), _p(hi, torch.bfloat16), _p(lo, torch.bfloat16, optional=True), x.numel(), _stream()))
    return hi, lo


def conv_forward_tc(s, x_hi, x_lo, w_hi, w_lo, addend=None, out=None, ci_real=None, bn_stats=None):
    """tcgen05 forward conv; x_* planes [n,t,h,w,ci] bf16, w_* planes [taps,co,ci] bf16 (lo planes None -> single-pass bf16)."""
    if out is None:
        out = torch.empty(s.n, s.to, s.ho, s.wo, s.co, dtype=torch.float32, device=x_hi.device)
    e0 = _t0()
    check(_lib.lib().avid_conv_forward_tc(C.byref(s), _p(x_hi, torch.bfloat16), _p(x_lo, torch.bfloat16, optional=True), _p(w_hi, torch.bfloat16),
                                          _p(w_lo, torch.bfloat16, optional=True), _p(addend, optional=True), _p(out),
                                          _p(bn_stats, torch.float64, optional=True), _stream()))
    if e0 is not None:      # bench.py roofline records, one family per kernel instantiation
        _t1(e0, "conv_pair_forward" if _lib.lib().avid_conv_tc_uses_cta_pairs(C.byref(s), 0) else "conv_forward_tc%d" % (128 if s.co % 128 == 0 else 64),
            _conv_flops(s, ci_real))
    return out


def bn_backward_sums(c, device):
    """(2, c) fp64 zeros for the BatchNorm-backward sums a fused input-gradient launch accumulates."""
    return _zeros((2, c), torch.float64, device)


def conv_dgrad_tc(s, d_hi, d_lo, w_hi, w_lo, addend=None, out=None, bn_fuse=None, addend_stride=None):
    """tcgen05 input gradient (any stride); d_* planes [n,to,ho,wo,co] bf16, w_* planes [taps,ci,co] bf16.
    bn_fuse = (z, BNState, gamma, beta, sums): also accumulate the BatchNorm-backward sums of the layer that produced this
    convolution's input (z: its conv output) -- see avid_bn_backward_fuse_t.
    addend_stride = (at, ah, aw): `addend` is the subsampled tensor [n, ceil(ti/at), ceil(hi/ah), ceil(wi/aw), ci] of
    avid_conv_dgrad_tc_sub (the input gradient of a strided 1x1x1 residual convolution, zero at every other pixel)."""
    if out is None:
        out = torch.empty(s.n, s.ti, s.hi, s.wi, s.ci, dtype=torch.float32, device=d_hi.device)
    sub = None
    if addend_stride is not None and addend is not None and tuple(addend_stride) != (1, 1, 1):
        at, ah, aw = addend_stride
        want = (s.n, -(-s.ti // at), -(-s.hi // ah), -(-s.wi // aw), s.ci)
        assert tuple(addend.shape) == want, (tuple(addend.shape), want)
        sub = (C.c_int32 * 3)(at, ah, aw)
    fuse = None
    if bn_fuse is not None:
        z, st, gamma, beta, sums = bn_fuse
        assert z.numel() == out.numel()
        fuse = _lib.BnBackwardFuse(_p(z), _p(st.mean), _p(st.invstd), _p(gamma), _p(beta), _p(sums, torch.float64))
    e0 = _t0()
    check(_lib.lib().avid_conv_dgrad_tc_sub(C.byref(s), _p(d_hi, torch.bfloat16), _p(d_lo, torch.bfloat16, optional=True), _p(w_hi, torch.bfloat16),
                                            _p(w_lo, torch.bfloat16, optional=True), _p(addend, optional=True), sub, _p(out),
                                            C.byref(fuse) if fuse is not None else None, _stream()))
    if e0 is not None:
        _t1(e0, "conv_pair_dgrad" if _lib.lib().avid_conv_tc_uses_cta_pairs(C.byref(s), 1) else "conv_dgrad_tc%d" % (128 if s.ci % 128 == 0 else 64), _conv_flops(s))
    return out


def conv_wgrad_tc(s, x_hi, x_lo, d_hi, d_lo, ci_real=None):
    """tcgen05 filter gradient; returns fp32 tap-major [taps, ci, co]."""
    dw = _zeros((s.kt * s.kh * s.kw, s.ci, s.co), torch.float32, x_hi.device)
    e0 = _t0()
    check(_lib.lib().avid_conv_wgrad_tc(C.byref(s), _p(x_hi, torch.bfloat16), _p(x_lo, torch.bfloat16, optional=True), _p(d_hi, torch.bfloat16),
                                        _p(d_lo, torch.bfloat16, optional=True), _p(dw), _stream()))
    _t1(e0, "conv_wgrad_tc", _conv_flops(s, ci_real))
    return dw


@_timed("stem_pack")
def stem_pack(x, wp, pad_left, need_lo=True):
    """(n, c, [t,] h, w) fp32 clip / spectrogram -> bf16 planes [n, t, h, wp, 4] of the stem kernels (c <= 4)."""
    n, c = x.shape[0], x.shape[1]
    t, h, w = (1,) * (5 - x.dim()) + tuple(x.shape[2:])
    hi = torch.empty(n, t, h, wp, 4, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty_like(hi) if need_lo else None
    check(_lib.lib().avid_stem_pack(_p(x), _p(hi, torch.bfloat16), _p(lo, torch.bfloat16, optional=True), n, c, t, h, w, wp, pad_left, _stream()))
    return hi, lo


def stem_filter_pack(w, need_lo=True):
    """PyTorch stem filter (co, ci, [kt,] kh, kw) -> bf16 planes [co, kt*kh, 32]."""
    co, ci = w.shape[0], w.shape[1]
    kt, kh, kw = (1,) * (5 - w.dim()) + tuple(w.shape[2:])
    hi = torch.empty(co, kt * kh, 32, dtype=torch.bfloat16, device=w.device)
    lo = torch.empty_like(hi) if need_lo else None
    check(_lib.lib().avid_stem_filter_pack(_p(w), _p(hi, torch.bfloat16), _p(lo, torch.bfloat16, optional=True), co, ci, kt, kh, kw, _stream()))
    return hi, lo


def stem_forward_tc(s, x_hi, x_lo, w_hi, w_lo, out=None, bn_stats=None):
    """tcgen05 stem convolution; s.ci is the real channel count, x_* the stem_pack planes, w_* the stem_filter_pack planes."""
    if out is None:
        out = torch.empty(s.n, s.to, s.ho, s.wo, s.co, dtype=torch.float32, device=x_hi.device)
    e0 = _t0()
    check(_lib.lib().avid_stem_forward_tc(C.byref(s), _p(x_hi, torch.bfloat16), _p(x_lo, torch.bfloat16, optional=True), x_hi.shape[3],
                                          _p(w_hi, torch.bfloat16), _p(w_lo, torch.bfloat16, optional=True), _p(out),
                                          _p(bn_stats, torch.float64, optional=True), _stream()))
    _t1(e0, "stem_forward_tc", _conv_flops(s))
    return out


def stem_wgrad_tc(s, x_hi, x_lo, d_hi, d_lo):
    """tcgen05 stem filter gradient -> fp32 tap-major [taps, 4, co]."""
    dw = _zeros((s.kt * s.kh * s.kw, 4, s.co), torch.float32, x_hi.device)
    e0 = _t0()
    check(_lib.lib().avid_stem_wgrad_tc(C.byref(s), _p(x_hi, torch.bfloat16), _p(x_lo, torch.bfloat16, optional=True), x_hi.shape[3],
                                        _p(d_hi, torch.bfloat16), _p(d_lo, torch.bfloat16, optional=True), _p(dw), _stream()))
    _t1(e0, "stem_wgrad_tc", _conv_flops(s))
    return dw


_plane_cache = {}        # weight.data_ptr() -> planes converted ahead of time by prepare_filter_planes (consumed by filter_to_planes)
_deferred_grads = None   # list of (dw_tap, out, co, ci, taps, ci_pad) while a tower's backward defers its filter-gradient layout conversion


def _ptr_array(tensors):
    return (C.c_void_p * len(tensors))(*[(t.data_ptr() if t is not None else None) for t in tensors])


@_timed("filter_to_planes")
def prepare_filter_planes(weights, need_lo=True):
    """Convert MANY PyTorch conv weights (co, ci, *k) to the tensor-core operand planes in ONE launch (avid_filter_to_planes_multi);
    the planes are handed out by the following filter_to_planes(w) calls.  One bf16 buffer backs all planes of the call."""
    global _plane_cache
    _plane_cache = {}
    weights = [w for w in weights if w is not None]
    if not weights:
        return
    dev = weights[0].device
    shapes = [(w.shape[0], w.shape[1], w[0, 0].numel()) for w in weights]
    per = 4 if need_lo else 2
    sizes = [co * ci * taps for co, ci, taps in shapes]
    buf = torch.empty(per * sum((n + 63) // 64 * 64 for n in sizes), dtype=torch.bfloat16, device=dev)
    f_hi, f_lo, d_hi, d_lo, off = [], [], [], [], 0
    for (co, ci, taps), n in zip(shapes, sizes):
        step = (n + 63) // 64 * 64
        views = [buf[off + k * step: off + k * step + n] for k in range(per)]
        off += per * step
        f_hi.append(views[0].view(taps, co, ci))
        d_hi.append(views[1].view(taps, ci, co))
        f_lo.append(views[2].view(taps, co, ci) if need_lo else None)
        d_lo.append(views[3].view(taps, ci, co) if need_lo else None)
    for w in weights:
        _p(w)
    ints = lambda vals: (C.c_int32 * len(vals))(*vals)
    check(_lib.lib().avid_filter_to_planes_multi(_ptr_array(weights), _ptr_array(f_hi), _ptr_array(f_lo) if need_lo else None, _ptr_array(d_hi),
                                                 _ptr_array(d_lo) if need_lo else None, ints([s_[0] for s_ in shapes]), ints([s_[1] for s_ in shapes]),
                                                 ints([s_[2] for s_ in shapes]), len(weights), _stream()))
    for w, a, b, c, d in zip(weights, f_hi, f_lo, d_hi, d_lo):
        _plane_cache[(w.data_ptr(), need_lo)] = ((a, b), (c, d))


def clear_filter_planes():
    global _plane_cache
    _plane_cache = {}


def defer_filter_gradients(on):
    """While on, filter_from_tapmajor only records its work; flush_filter_gradients converts every recorded gradient in ONE launch."""
    global _deferred_grads
    _deferred_grads = [] if on else None


@_timed("filter_from_tap")
def flush_filter_gradients():
    global _deferred_grads
    todo, _deferred_grads = _deferred_grads or [], None
    if not todo:
        return
    ints = lambda k: (C.c_int32 * len(todo))(*[t[k] for t in todo])
    check(_lib.lib().avid_filter_from_tapmajor_multi(_ptr_array([t[0] for t in todo]), _ptr_array([t[1] for t in todo]), ints(2), ints(3), ints(4),
                                                     ints(5), len(todo), _stream()))


@_timed("filter_to_planes")
def filter_to_planes(w, need_lo=True):
    """PyTorch conv weight (co, ci, *k) -> bf16 planes ((fwd_hi, fwd_lo) [taps, co, ci], (dgrad_hi, dgrad_lo) [taps, ci, co])."""
    hit = _plane_cache.pop((w.data_ptr(), need_lo), None)
    if hit is not None:
        return hit
    co, ci = w.shape[0], w.shape[1]
    taps = w[0, 0].numel()
    mk = lambda *shape: torch.empty(shape, dtype=torch.bfloat16, device=w.device)
    f_hi, d_hi = mk(taps, co, ci), mk(taps, ci, co)
    f_lo, d_lo = (mk(taps, co, ci), mk(taps, ci, co)) if need_lo else (None, None)
    check(_lib.lib().avid_filter_to_planes(_p(w), _p(f_hi, torch.bfloat16), _p(f_lo, torch.bfloat16, optional=True), _p(d_hi, torch.bfloat16),
                                           _p(d_lo, torch.bfloat16, optional=True), co, ci, taps, _stream()))
    return (f_hi, f_lo), (d_hi, d_lo)


def filter_to_tapmajor(w, ci_pad=None, transpose=True):
    """PyTorch conv weight (co, ci, *k) -> ([taps, ci_pad, co], [taps, co, ci_pad] or None)."""
    co, ci = w.shape[0], w.shape[1]
    taps = w[0, 0].numel()
    ci_pad = ci if ci_pad is None else ci_pad
    w_tap = torch.empty(taps, ci_pad, co, dtype=torch.float32, device=w.device)
    w_tap_t = torch.empty(taps, co, ci_pad, dtype=torch.float32, device=w.device) if transpose else None
    check(_lib.lib().avid_filter_to_tapmajor(_p(w), _p(w_tap), _p(w_tap_t, optional=True), co, ci, taps, ci_pad, _stream()))
    return w_tap, w_tap_t


@_timed("filter_from_tap")
def filter_from_tapmajor(dw_tap, like):
    co, ci = like.shape[0], like.shape[1]
    taps, ci_pad = dw_tap.shape[0], dw_tap.shape[1]
    out = torch.empty_like(like)
    if _deferred_grads is not None:
        _p(dw_tap)
        _deferred_grads.append((dw_tap, out, co, ci, taps, ci_pad))      # dw_tap stays alive until the flush
        return out
    check(_lib.lib().avid_filter_from_tapmajor(_p(dw_tap), _p(out), co, ci, taps, ci_pad, _stream()))
    return out


@_timed("layout")
def nchw_to_nhwc(x, c_pad=None):
    """(n, c, *spatial) -> (n, *spatial, c_pad)."""
    n, c = x.shape[0], x.shape[1]
    sp = tuple(x.shape[2:])
    thw = x[0, 0].numel()
    c_pad = c if c_pad is None else c_pad
    out = torch.empty((n,) + sp + (c_pad,), dtype=torch.float32, device=x.device)
    check(_lib.lib().avid_nchw_to_nhwc(_p(x), _p(out), n, c, thw, c_pad, _stream()))
    return out


def nhwc_to_nchw(x):
    """(n, *spatial, c) -> (n, c, *spatial)."""
    n, c = x.shape[0], x.shape[-1]
    sp = tuple(x.shape[1:-1])
    out = torch.empty((n, c) + sp, dtype=torch.float32, device=x.device)
    check(_lib.lib().avid_nhwc_to_nchw(_p(x), _p(out), n, c, x[0, ..., 0].numel(), _stream()))
    return out


class BNState:
    """Per-call buffers of one train-mode BatchNorm: statistics, saved mean/invstd, folded scale/shift."""
    __slots__ = ("stats", "mean", "invstd", "scale", "shift", "frozen")

    def __init__(self, c, device):
        self.frozen = False          # True: the layer ran with running statistics (model.eval()): its backward is affine-only
        self.stats = _zeros((2, c), torch.float64, device)
        buf = torch.empty(4, c, dtype=torch.float32, device=device)
        self.mean, self.invstd, self.scale, self.shift = buf[0], buf[1], buf[2], buf[3]


@_timed("bn_finalize")
def bn_train_stats(x, gamma, beta, running_mean, running_var, eps=BN_EPS, momentum=BN_MOMENTUM, state=None):
    """x (..., c) channels-last.  Computes batch statistics, updates running stats, returns BNState.  With `state` given its
    `stats` were already accumulated by the producing convolution's epilogue and only the finalize kernel runs."""
    c = x.shape[-1]
    rows = x.numel() // c
    L = _lib.lib()
    s = state
    if s is None:
        s = BNState(c, x.device)
        check(L.avid_bn_stats(_p(x), rows, c, _p(s.stats, torch.float64), _stream()))
    check(L.avid_bn_finalize(_p(s.stats, torch.float64), rows, c, _p(gamma), _p(beta), eps, momentum,
                             _p(running_mean, optional=True), _p(running_var, optional=True),
                             _p(s.mean), _p(s.invstd), _p(s.scale), _p(s.shift), _stream()))
    return s


class Act:
    """A channels-last activation in the representations the kernels consume: fp32 and / or the bf16
    planes (hi, lo) with x ~= hi + lo that feed the tcgen05 convolutions."""
    __slots__ = ("f32", "hi", "lo")

    def __init__(self, f32=None, hi=None, lo=None):
        self.f32, self.hi, self.lo = f32, hi, lo

    @property
    def shape(self):
        return (self.f32 if self.f32 is not None else self.hi).shape

    @property
    def device(self):
        return (self.f32 if self.f32 is not None else self.hi).device

    def ensure_planes(self, x3):
        if self.hi is None or (x3 and self.lo is None):
            self.hi, self.lo = split_bf16(self.f32, x3)
        return self


def bn_relu_forward(x, scale, shift, out=None):
    c = x.shape[-1]
    if out is None:
        out = torch.empty_like(x)
    check(_lib.lib().avid_bn_relu_forward(_p(x), _p(scale), _p(shift), _p(out), x.numel() // c, c, _stream()))
    return out


@_timed("bn_relu_fwd")
def bn_relu_forward_act(x, scale, shift, want_f32, want_planes, x3):
    """relu(x * scale + shift) as an Act with the requested representations, in one pass over x."""
    c = x.shape[-1]
    y = torch.empty_like(x) if want_f32 else None
    hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if want_planes else None
    lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if want_planes and x3 else None
    check(_lib.lib().avid_bn_relu_forward_ex(_p(x), _p(scale), _p(shift), _p(y, optional=True), _p(hi, torch.bfloat16, optional=True),
                                             _p(lo, torch.bfloat16, optional=True), x.numel() // c, c, _stream()))
    return Act(y, hi, lo)


def bn_relu_backward(x, dy, s, gamma, beta, dx=None):
    """Backward of y = relu(bn_train(x)).  Returns (dx, dgamma, dbeta)."""
    act, dgamma, dbeta = bn_relu_backward_act(x, dy, s, gamma, beta, True, False, False, dx=dx)
    return act.f32, dgamma, dbeta


@_timed("bn_relu_bwd")
def bn_relu_backward_act(x, dy, s, gamma, beta, want_f32, want_planes, x3, dx=None, sums=None):
    """Backward of y = relu(bn_train(x)); the gradient w.r.t. x as an Act (fp32 and / or bf16 planes).  `sums`: the reduction
    was already done by the epilogue of the input-gradient launch that produced dy (conv_dgrad_tc(bn_fuse=...))."""
    c = x.shape[-1]
    rows = x.numel() // c
    have_sums = sums is not None
    if not have_sums:
        sums = _zeros((2, c), torch.float64, x.device)
    if want_f32 and dx is None:
        dx = torch.empty_like(x)
    hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if want_planes else None
    lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if want_planes and x3 else None
    dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(beta)
    L = _lib.lib()
    if not have_sums:
        check(L.avid_bn_relu_backward_reduce(_p(x), _p(dy), _p(s.mean), _p(s.invstd), _p(gamma), _p(beta), rows, c, _p(sums, torch.float64), _stream()))
    sums, dgamma, dbeta = _frozen_sums(s, sums, dgamma, dbeta)
    check(L.avid_bn_relu_backward_apply_ex(_p(x), _p(dy), _p(s.mean), _p(s.invstd), _p(gamma), _p(beta), _p(sums, torch.float64), rows, c,
                                           _p(dx, optional=True), _p(hi, torch.bfloat16, optional=True), _p(lo, torch.bfloat16, optional=True),
                                           _p(dgamma) if not s.frozen else None, _p(dbeta) if not s.frozen else None, _stream()))
    return Act(dx if want_f32 else None, hi, lo), dgamma, dbeta


def _frozen_sums(s, sums, dgamma, dbeta):
    """A layer that ran with its RUNNING statistics (model.eval(), frozen-BatchNorm fine-tuning) is a per-channel affine map: its
    input gradient is dy * relu' * gamma * invstd without the batch-statistic terms, while dgamma / dbeta are still the reduced
    sums.  The apply kernels compute k * (g - sums[0]/n - xhat * sums[1]/n): hand them zero sums."""
    if not s.frozen:
        return sums, dgamma, dbeta
    return torch.zeros_like(sums), sums[1].float(), sums[0].float()


@_timed("bn_pool_fwd")
def bn_relu_maxpool_forward(z, scale, shift, want_f32, want_planes, x3):
    """maxpool_1x3x3(relu(z * scale + shift)) without materialising the ReLU output.  z (n, t, h, w, c).
    Returns (Act of the pooled tensor (n, t, ho, wo, c), argmax uint8)."""
    n, t, h, w, c = z.shape
    ho, wo = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
    shape = (n, t, ho, wo, c)
    p = torch.empty(shape, dtype=torch.float32, device=z.device) if want_f32 else None
    hi = torch.empty(shape, dtype=torch.bfloat16, device=z.device) if want_planes else None
    lo = torch.empty(shape, dtype=torch.bfloat16, device=z.device) if want_planes and x3 else None
    am = torch.empty(shape, dtype=torch.uint8, device=z.device)
    check(_lib.lib().avid_bn_relu_maxpool_forward(_p(z), _p(scale), _p(shift), _p(p, optional=True), _p(hi, torch.bfloat16, optional=True),
                                                  _p(lo, torch.bfloat16, optional=True), _p(am, torch.uint8), n * t, h, w, c, ho, wo, _stream()))
    return Act(p, hi, lo), am


@_timed("bn_pool_bwd")
def bn_relu_maxpool_backward_act(z, pooled, argmax, dyp, s, gamma, beta, want_f32, want_planes, x3):
    """Backward of pooled = maxpool(relu(bn_train(z))): the gradient w.r.t. z as an Act, dgamma, dbeta."""
    n, t, h, w, c = z.shape
    ho, wo = argmax.shape[2], argmax.shape[3]
    sums = _zeros((2, c), torch.float64, z.device)
    dz = torch.empty_like(z) if want_f32 else None
    hi = torch.empty(z.shape, dtype=torch.bfloat16, device=z.device) if want_planes else None
    lo = torch.empty(z.shape, dtype=torch.bfloat16, device=z.device) if want_planes and x3 else None
    dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(beta)
    L = _lib.lib()
    common = (_p(z), _p(argmax, torch.uint8), _p(dyp), _p(s.mean), _p(s.invstd), _p(gamma), _p(beta))
    check(L.avid_bn_relu_maxpool_backward_reduce(_p(z), _p(pooled), *common[1:], n * t, h, w, c, ho, wo, _p(sums, torch.float64), _stream()))
    sums, dgamma, dbeta = _frozen_sums(s, sums, dgamma, dbeta)
    check(L.avid_bn_relu_maxpool_backward_apply(*common, _p(sums, torch.float64), n * t, h, w, c, ho, wo, _p(dz, optional=True),
                                                _p(hi, torch.bfloat16, optional=True), _p(lo, torch.bfloat16, optional=True),
                                                _p(dgamma) if not s.frozen else None, _p(dbeta) if not s.frozen else None, _stream()))
    return Act(dz, hi, lo), dgamma, dbeta


def maxpool_1x3x3_forward(x, need_argmax=True):
    """x (n, t, h, w, c) -> (y (n, t, ho, wo, c), argmax uint8 (n, t, ho, wo, c) or None)."""
    n, t, h, w, c = x.shape
    ho, wo = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
    y = torch.empty(n, t, ho, wo, c, dtype=torch.float32, device=x.device)
    am = torch.empty(n, t, ho, wo, c, dtype=torch.uint8, device=x.device) if need_argmax else None
    check(_lib.lib().avid_maxpool_1x3x3_forward(_p(x), _p(y), _p(am, torch.uint8, optional=True), n * t, h, w, c, ho, wo, _stream()))
    return y, am


def maxpool_1x3x3_backward(argmax, dy, in_shape):
    n, t, h, w, c = in_shape
    dx = torch.empty(in_shape, dtype=torch.float32, device=dy.device)
    check(_lib.lib().avid_maxpool_1x3x3_backward(_p(argmax, torch.uint8), _p(dy), _p(dx), n * t, h, w, c, argmax.shape[2], argmax.shape[3], _stream()))
    return dx


@_timed("global_pool")
def global_maxpool_forward(x):
    """x (n, ..., c) -> (y (n, c), argmax (n, c) int32)."""
    n, c = x.shape[0], x.shape[-1]
    thw = x.numel() // (n * c)
    y = torch.empty(n, c, dtype=torch.float32, device=x.device)
    am = torch.empty(n, c, dtype=torch.int32, device=x.device)
    check(_lib.lib().avid_global_maxpool_forward(_p(x), _p(y), _p(am, torch.int32), n, thw, c, _stream()))
    return y, am


@_timed("global_pool")
def global_maxpool_backward(dy, argmax, shape):
    dx = torch.zeros(shape, dtype=torch.float32, device=dy.device)
    n, c = dy.shape
    check(_lib.lib().avid_global_maxpool_backward(_p(dy), _p(argmax, torch.int32), _p(dx), n, dx.numel() // (n * c), c, _stream()))
    return dx


@_timed("linear")
def linear_forward(x, w, b, relu):
    rows, in_f = x.shape
    out_f = w.shape[0]
    y = torch.empty(rows, out_f, dtype=torch.float32, device=x.device)
    check(_lib.lib().avid_linear_forward(_p(x), _p(w), _p(b, optional=True), _p(y), rows, in_f, out_f, int(relu), _stream()))
    return y


@_timed("linear")
def linear_backward(x, w, y, dy, relu, need_dx=True):
    """dy is modified in place when relu.  Returns (dx or None, dw, db)."""
    rows, in_f = x.shape
    out_f = w.shape[0]
    dx = torch.empty_like(x) if need_dx else None
    dw, db = torch.empty_like(w), torch.empty(out_f, dtype=torch.float32, device=x.device)
    check(_lib.lib().avid_linear_backward(_p(x), _p(w), _p(y, optional=True), _p(dy), _p(dx, optional=True), _p(dw), _p(db),
                                          rows, in_f, out_f, int(relu), _stream()))
    return dx, dw, db


def add_(a, b):
    check(_lib.lib().avid_add_inplace(_p(a), _p(b), a.numel(), _stream()))
    return a


@_timed("adam")
def adam_step_multi_(params, grads, exp_avgs, exp_avg_sqs, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0):
    """One Adam update of many fp32 tensors (32 per kernel launch)."""
    n = len(params)
    if n == 0:
        return
    for t in list(params) + list(grads) + list(exp_avgs) + list(exp_avg_sqs):
        _p(t)                                      # validates device / dtype / contiguity
    arr = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])
    sizes = (C.c_int64 * n)(*[t.numel() for t in params])
    check(_lib.lib().avid_adam_step_multi(arr(params), arr(grads), arr(exp_avgs), arr(exp_avg_sqs), sizes, n, step, lr, betas[0], betas[1], eps,
                                          weight_decay, grad_scale, _stream()))


def peer_ptrs(ptrs):
    """avid_peer_ptrs_t from a list of raw device pointers (symmetric-memory handle.buffer_ptrs: rank r's buffer as mapped here)."""
    if len(ptrs) > _lib.AVID_MAX_PEERS:
        raise ValueError("at most %d peers" % _lib.AVID_MAX_PEERS)
    s = _lib.PeerPtrs()
    for i, p in enumerate(ptrs):
        s.ptr[i] = int(p)
    return s


@_timed("adam")
def adam_shard_step_(flat_param, peer_grads, world, exp_avg, exp_avg_sq, begin, count, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                     grad_scale=1.0):
    """Fused reduce-scatter (P2P loads from every rank's flat gradients) + Adam on the shard [begin, begin + count) this rank owns."""
    check(_lib.lib().avid_adam_shard_step(_p(flat_param), C.byref(peer_grads), world, _p(exp_avg), _p(exp_avg_sq), begin, count, step, lr,
                                          betas[0], betas[1], eps, weight_decay, grad_scale, _stream()))


@_timed("adam")
def pull_shards_(flat_param, peer_params, world, rank, shard):
    """All-gather of the updated parameter shards by P2P loads into the local flat parameter buffer."""
    check(_lib.lib().avid_pull_shards(_p(flat_param), C.byref(peer_params), world, rank, shard, _stream()))


def adam_step_(param, grad, exp_avg, exp_avg_sq, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0):
    check(_lib.lib().avid_adam_step(_p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), param.numel(), step, lr, betas[0], betas[1], eps,
                                    weight_decay, grad_scale, _stream()))
