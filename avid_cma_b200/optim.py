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
