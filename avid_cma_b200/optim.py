"""Adam on the libavid_b200 kernel: the reference builds torch.optim.Adam(lr, weight_decay, betas)
(utils/main_utils.py:250-256); this optimizer has the same update rule and state_dict layout
(`exp_avg`, `exp_avg_sq`, `step` per parameter) but applies it with avid_adam_step."""
import torch

from . import ops


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            by_step = {}
            for p in group['params']:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                # torch.optim.Adam state_dicts store `step` as a tensor: normalise to a python int (dict key, kernel argument)
                st['step'] = int(st['step']) + 1
                by_step.setdefault(st['step'], []).append((p.data, p.grad.contiguous(), st['exp_avg'], st['exp_avg_sq']))
            for step, items in by_step.items():      # normally one entry: every parameter has taken the same number of steps
                ps, gs, ms, vs = zip(*items)
                ops.adam_step_multi_(ps, gs, ms, vs, step, group['lr'], group['betas'], group['eps'], group['weight_decay'], grad_scale)
        return loss


class ShardedAdam(torch.optim.Optimizer):
    """Data-parallel Adam whose gradient exchange is fused into the update, over NVLink peer memory (csrc/shard_optim.cu).

    Replaces the pair the reference uses for W > 1 ranks -- DistributedDataParallel's bucketed NCCL all-reduce of the gradients
    (utils/main_utils.py:105-117) and torch.optim.Adam.step on every rank (main_utils.py:250-256) -- with the SAME arithmetic
    (average of the ranks' gradients in rank order, then Adam with L2 weight decay):

      * the parameters of the model become views of ONE flat symmetric buffer (same layout on every rank, peer-mapped), the
        gradients of a step are copied into a second one;
      * rank r owns the shard [r * S, (r + 1) * S): `avid_adam_shard_step` reads that shard of every rank's gradients by P2P loads,
        reduces, and updates its shard of the parameters and of the moments (which only the owner keeps: 1 / W of the optimizer
        state per GPU); `avid_pull_shards` then copies the other ranks' updated shards into the local flat parameters;
      * two symmetric-memory barriers per step order the ranks; no collective library call, no SM carve-out, ~0.2 ms per step.

    Wrap the model in `LocalGradients` instead of DistributedDataParallel (the gradients stay local until step()).  Needs all
    ranks on one node with peer access (torch.distributed._symmetric_memory); `available()` tells.  state_dict() / load_state_dict()
    keep torch.optim.Adam's layout -- they gather / slice the moment shards, so EVERY rank must call them (like the row-sharded
    memory bank, utils/main_utils.py::reference_state_dict)."""

    collective_state_dict = True

    @staticmethod
    def available():
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and dist.get_backend() == 'nccl'):
            return False
        try:
            import torch.distributed._symmetric_memory  # noqa: F401
        except Exception:
            return False
        return dist.get_world_size() <= 16

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise ValueError('ShardedAdam takes a single parameter group')
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        ps = [p for p in self.param_groups[0]['params']]
        if not ps or any((not p.is_cuda) or p.dtype != torch.float32 for p in ps):
            raise ValueError('ShardedAdam needs fp32 CUDA parameters')
        dev = ps[0].device
        # flat layout: every parameter starts at a multiple of 4 elements; the total is padded to world * shard
        self._offsets, off = [], 0
        for p in ps:
            self._offsets.append(off)
            off += (p.numel() + 3) // 4 * 4
        self.total = off
        self.shard = ((off + self.world - 1) // self.world + 3) // 4 * 4
        n = self.world * self.shard
        self.flat_param = symm.empty(n, dtype=torch.float32, device=dev)
        self.flat_grad = symm.empty(n, dtype=torch.float32, device=dev)
        self.flat_param.zero_()
        self.flat_grad.zero_()
        self._hp = symm.rendezvous(self.flat_param, self.group)
        self._hg = symm.rendezvous(self.flat_grad, self.group)
        with torch.no_grad():
            for p, o in zip(ps, self._offsets):
                self.flat_param[o:o + p.numel()].copy_(p.detach().reshape(-1))
            dist.broadcast(self.flat_param, dist.get_global_rank(self.group, 0) if hasattr(dist, 'get_global_rank') else 0, group=self.group)   # DDP's initial sync
            for p, o in zip(ps, self._offsets):
                p.data = self.flat_param[o:o + p.numel()].view(p.shape)      # the model now reads the flat buffer
        self._grad_views = [self.flat_grad[o:o + p.numel()].view(p.shape) for p, o in zip(ps, self._offsets)]
        self._peer_param = ops.peer_ptrs(self._hp.buffer_ptrs)
        self._peer_grad = ops.peer_ptrs(self._hg.buffer_ptrs)
        self.exp_avg = torch.zeros(self.shard, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(self.shard, dtype=torch.float32, device=dev)
        self._step = 0
        torch.cuda.synchronize(dev)
        self._hp.barrier(channel=0)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        g = self.param_groups[0]
        ps = g['params']
        have = [(v, p.grad) for v, p in zip(self._grad_views, ps) if p.grad is not None]
        missing = [v for v, p in zip(self._grad_views, ps) if p.grad is None]
        if have:
            torch._foreach_copy_([v for v, _ in have], [gr for _, gr in have])
        if missing:
            torch._foreach_zero_(missing)
        self._step += 1
        self._hg.barrier(channel=0)          # every rank's gradients are in its flat buffer
        ops.adam_shard_step_(self.flat_param, self._peer_grad, self.world, self.exp_avg, self.exp_avg_sq, self.rank * self.shard, self.shard,
                             self._step, g['lr'], g['betas'], g['eps'], g['weight_decay'], 1.0 / self.world)
        self._hp.barrier(channel=1)          # every shard is updated (and every rank is done reading gradients)
        ops.pull_shards_(self.flat_param, self._peer_param, self.world, self.rank, self.shard)
        return loss

    # ---- torch.optim.Adam's state_dict layout (COLLECTIVE: the moment shards are gathered / sliced) ----
    def _gather(self, shard):
        import torch.distributed as dist
        full = torch.empty(self.world * self.shard, dtype=torch.float32, device=shard.device)
        dist.all_gather_into_tensor(full, shard, group=self.group)
        return full

    def state_dict(self):
        m, v = self._gather(self.exp_avg), self._gather(self.exp_avg_sq)
        ps = self.param_groups[0]['params']
        state = {}
        if self._step > 0:
            for i, (p, o) in enumerate(zip(ps, self._offsets)):
                state[i] = {'step': torch.tensor(float(self._step)), 'exp_avg': m[o:o + p.numel()].view(p.shape).clone(),
                            'exp_avg_sq': v[o:o + p.numel()].view(p.shape).clone()}
        group = {k: val for k, val in self.param_groups[0].items() if k != 'params'}
        group['params'] = list(range(len(ps)))
        return {'state': state, 'param_groups': [group]}

    def load_state_dict(self, state_dict):
        ps = self.param_groups[0]['params']
        group = state_dict['param_groups'][0]
        for k, val in group.items():
            if k != 'params':
                self.param_groups[0][k] = val
        m = torch.zeros(self.world * self.shard, dtype=torch.float32, device=self.exp_avg.device)
        v = torch.zeros_like(m)
        step = 0
        for i, (p, o) in enumerate(zip(ps, self._offsets)):
            st = state_dict['state'].get(i, state_dict['state'].get(str(i)))
            if st is None:
                continue
            step = max(step, int(st['step']))
            m[o:o + p.numel()].copy_(st['exp_avg'].reshape(-1))
            v[o:o + p.numel()].copy_(st['exp_avg_sq'].reshape(-1))
        lo = self.rank * self.shard
        self.exp_avg.copy_(m[lo:lo + self.shard])
        self.exp_avg_sq.copy_(v[lo:lo + self.shard])
        self._step = step


class LocalGradients(torch.nn.Module):
    """What DistributedDataParallel is to torch.optim.Adam, this is to ShardedAdam: the wrapper main-avid.py:100 expects
    (`model.module`, `module.`-prefixed state_dict keys), WITHOUT a gradient all-reduce -- the exchange happens inside
    ShardedAdam.step().  Buffers (BatchNorm running statistics) are made identical once, like DDP's constructor does."""

    def __init__(self, module, group=None):
        import torch.distributed as dist
        super().__init__()
        self.module = module
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            for b in module.buffers():
                dist.broadcast(b, 0, group=group)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)
