"""NCECriterion state holder (reference: criterions/nce.py:14-58).

In this implementation the NCE arithmetic lives inside the fused CUDA kernel
(csrc/nce.cu, avid_nce_forward_backward); this module only owns the frozen partition
constant `avg_exp_score` so that state_dicts stay interchangeable with the reference
(key `criterion.avg_exp_score`).
"""
import torch
from torch import nn


class NCECriterion(nn.Module):
    def __init__(self, nLem):
        super().__init__()
        self.nLem = nLem
        self.register_buffer('avg_exp_score', torch.tensor(-1.))
        self._z_ready = None   # unknown until checked once on the host (or after load_state_dict)
        self._register_load_state_dict_pre_hook(self._on_load)

    def _on_load(self, state_dict, prefix, *args):
        self._z_ready = None
        key = prefix + 'avg_exp_score'
        if key in state_dict and state_dict[key].dim() != 0:   # reference saves shape [1] in single-process runs
            state_dict[key] = state_dict[key].reshape(())

    def z_ready(self):
        """True once the partition function has been estimated (nce.py:21-24).  One host sync, then cached."""
        if self._z_ready is None:
            self._z_ready = bool(float(self.avg_exp_score) > 0)
        return self._z_ready

    def mark_ready(self):
        self._z_ready = True
