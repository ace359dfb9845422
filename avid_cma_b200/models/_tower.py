"""autograd plumbing shared by the two encoders: one autograd.Function per tower whose forward/backward
are straight-line sequences of libavid_b200 kernel launches (no per-op autograd nodes)."""
import os

import torch

from .. import ops

_MATH = {'fp32': ops.MATH_FP32, 'bf16x3': ops.MATH_BF16X3, 'bf16': ops.MATH_BF16}


def default_math():
    return _MATH[os.environ.get('AVID_MATH', 'bf16x3')]


def prepare_tower_filters(tower):
    """All tensor-core filters of the tower -> bf16 operand planes in one launch (instead of one launch per layer)."""
    if tower.math == ops.MATH_FP32:
        return
    ws = [m.weight.detach() for m in tower.modules()
          if isinstance(m, (torch.nn.Conv2d, torch.nn.Conv3d)) and m.in_channels % 64 == 0 and m.out_channels % 64 == 0]
    ops.prepare_filter_planes(ws, need_lo=tower.math == ops.MATH_BF16X3)


def _run_forward(tower, x):
    arena = tower.__dict__.setdefault('_fwd_arena', ops.ZeroArena())
    arena.begin(x.device)
    ops.set_arena(arena)
    try:
        prepare_tower_filters(tower)
        return tower._fwd(x, tower.training, tower.math)
    finally:
        ops.set_arena(None)
        ops.clear_filter_planes()


def _run_backward(tower, dpooled, saved):
    grads = {}
    arena = tower.__dict__.setdefault('_bwd_arena', ops.ZeroArena())
    arena.begin(dpooled.device)
    ops.set_arena(arena)
    ops.defer_filter_gradients(True)          # every filter gradient is brought to the PyTorch layout by ONE launch at the end
    try:
        tower._bwd(dpooled, saved, grads, tower.math)
        ops.side_join()                       # filter gradients enqueued on the side stream (AVID_WGRAD_STREAM=1)
        ops.flush_filter_gradients()
    finally:
        ops.defer_filter_gradients(False)
        ops.set_arena(None)
    return grads


class GraphedTower:
    """CUDA graphs of one tower's training forward and backward for one input shape.

    A tower step is ~150 (forward) + ~170 (backward) dependent kernel launches of 5 us .. 2 ms: replaying them from two captured
    graphs removes the launch gaps and the host work per launch (ctypes call, tensor-map encoding, allocator).  Everything a
    launch reads or writes has a fixed address: the input is copied into a static buffer, activations / saved tensors / gradients
    live in the graph's private memory pool, parameters and BatchNorm buffers are updated in place by the optimizer.  The first
    calls run eagerly (they also size the zero arenas); the capture happens on call number `kWarmup + 1`."""
    kWarmup = 2

    def __init__(self, tower, x):
        self.tower = tower
        self.x = torch.empty_like(x)
        self.fwd = self.bwd = None
        self.pooled = self.saved = self.dpooled = self.grads = None
        self.launches_fwd = self.launches_bwd = 0
        self.pending = False        # a forward whose backward has not run yet owns the static buffers

    def forward(self, x):
        self.x.copy_(x)
        self.pending = True
        if self.fwd is None:
            torch.cuda.synchronize()
            n0 = ops.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.pooled, self.saved = _run_forward(self.tower, self.x)
            self.launches_fwd = ops.launch_count() - n0     # counted by the library while being captured: the first replay runs them
            self.fwd = g
            self.fwd.replay()
            return self.pooled
        self.fwd.replay()
        ops.add_launches(self.launches_fwd)
        return self.pooled

    def backward(self, dpooled):
        if self.bwd is None:
            self.dpooled = torch.empty_like(dpooled)
            self.dpooled.copy_(dpooled)
            torch.cuda.synchronize()
            n0 = ops.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=self.fwd.pool()):
                self.grads = _run_backward(self.tower, self.dpooled, self.saved)
            self.launches_bwd = ops.launch_count() - n0
            self.bwd = g
        else:
            self.dpooled.copy_(dpooled)
            ops.add_launches(self.launches_bwd)
        self.bwd.replay()
        self.pending = False
        return self.grads


def _graphs_enabled():
    return os.environ.get('AVID_CUDA_GRAPH', '1') == '1' and not ops.profiling()


class TowerFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tower, x, *params):
        x = x.detach().contiguous().float()
        graphed = None
        if tower.training and _graphs_enabled():
            # one graph pair per (input shape, arithmetic, parameter storage): anything else falls back to eager launches
            key = (tuple(x.shape), tower.math, x.device.index, params[0].data_ptr(), params[-1].data_ptr())
            cache = tower.__dict__.setdefault('_graphs', {})
            entry = cache.get(key)
            if entry is None:
                entry = cache[key] = [0, None]
            entry[0] += 1
            if entry[0] > GraphedTower.kWarmup and entry[1] is not False and not (entry[1] is not None and entry[1].pending):
                if entry[1] is None:
                    entry[1] = GraphedTower(tower, x)
                graphed = entry[1]
                try:
                    pooled = graphed.forward(x)
                except Exception as e:        # noqa: BLE001 -- capture is an optimisation: fall back to eager launches, loudly
                    import warnings
                    warnings.warn('avid_cma_b200: CUDA-graph capture of %s failed (%r); running eager' % (type(tower).__name__, e))
                    entry[1] = False
                    graphed = None
        if graphed is None:
            pooled, saved = _run_forward(tower, x)
            ctx.saved = saved
        ctx.tower, ctx.graphed, ctx.params = tower, graphed, params
        return pooled

    @staticmethod
    def backward(ctx, dpooled):
        dpooled = dpooled.contiguous()
        if ctx.graphed is not None:
            grads = ctx.graphed.backward(dpooled)
            # the gradient tensors are the graph's static buffers: hand out copies if a parameter still holds last step's gradient
            # (autograd would add the buffer to itself), otherwise autograd just takes the reference
            if any(p.grad is not None for p in ctx.params):
                grads = {k: v.clone() for k, v in grads.items()}
        else:
            grads = _run_backward(ctx.tower, dpooled, ctx.saved)
            ctx.saved = None
        return (None, None) + tuple(grads.get(p) for p in ctx.params)


class TowerMixin:
    """forward(x, return_embs=False) of models/video.py:44-54 / models/audio.py:34-44."""
    math = None

    def forward(self, x, return_embs=False):
        if self.math is None:
            self.math = default_math()
        if not x.is_cuda:
            raise RuntimeError('avid_cma_b200 encoders run on CUDA tensors only (no CPU fallback)')
        if getattr(self, '_is_replica', False):
            # nn.DataParallel over several GPUs runs its replicas in threads of ONE process; the launch bookkeeping of this package
            # (zero arena, deferred layout conversions) is per process by design: one process per GPU
            raise RuntimeError('avid_cma_b200 runs one process per GPU: use --multiprocessing-distributed / torchrun (DistributedDataParallel) '
                               'instead of multi-GPU nn.DataParallel')
        if return_embs:
            # feature taps for evaluation (utils/eval_utils.py:208,324,343): forward only, reference NC(D)HW layout
            if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and self.training:
                raise RuntimeError('return_embs=True is supported for inference (model.eval() / torch.no_grad()) only')
            taps = {}
            with torch.no_grad():
                pooled, _ = self._fwd(x.contiguous().float(), self.training, self.math, taps=taps)
            two_d = x.dim() == 4
            out = {}
            for k, v in taps.items():
                v = ops.nhwc_to_nchw(v.contiguous())
                out[k] = v.squeeze(2) if two_d else v
            out['pool'] = pooled.view(pooled.shape + ((1, 1) if two_d else (1, 1, 1)))
            return out
        params = tuple(p for p in self.parameters())
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            pooled = TowerFunction.apply(self, x, *params)
        else:
            pooled, _ = self._fwd(x.contiguous().float(), self.training, self.math)
        return pooled.view(pooled.shape + ((1, 1) if x.dim() == 4 else (1, 1, 1)))
