"""Oracle training step (TEST INFRASTRUCTURE, see oracle/__init__.py): the reference's run_phase body
(main-avid.py:155-184) -- forward of both towers, AVID criterion, backward, Adam -- restated on torch CPU
with the oracle towers / criterion.  Used as the checker in tests and smoke(), and timed by bench.py's
`cpu_baseline` / `--impl reference` legs as the reference's CPU path."""
import torch

from . import criterion as oc
from . import synth, towers


class OracleTrainer:
    def __init__(self, num_data, num_negatives=1024, momentum=0.5, lr=2e-4, weight_decay=1e-5, seed=0, keys=None, device="cpu"):
        """device: "cpu" (the checker / the reference's CPU path) or a CUDA device -- the same stock torch ops then run on
        cuDNN / cuBLAS / ATen, which is how the reference itself executes on a GPU (bench.py --impl torch_gpu, diagnostic arm)."""
        self.sd = {k: v.to(device) for k, v in synth.fill_state_dict(towers.state_dict_template(), seed=seed).items()}
        self.params = [self.sd[k].requires_grad_(True) for k in towers.param_keys(self.sd)]
        self.opt = torch.optim.Adam(self.params, lr=lr, weight_decay=weight_decay, betas=(0.9, 0.999))   # main_utils.py:250-256
        self.bank_v = synth.bank(num_data, seed=seed, tag="bank_v").to(device)
        self.bank_a = synth.bank(num_data, seed=seed, tag="bank_a").to(device)
        self.N, self.K, self.momentum = num_data, num_negatives, momentum
        self.keys = keys or oc.avid_keys(num_negatives)
        self.Z = -1.0
        self.gen = torch.Generator().manual_seed(seed)

    def step(self, video, audio, y, neg_idx=None):
        """One optimisation step; returns the python float loss (the `.item()` of main-avid.py:174)."""
        if neg_idx is None:   # avid.py:82-86
            # drawn on the host RNG and copied to the device, like alias_method.py:64-71 + avid.py:84
            raw = torch.randint(0, self.N - 1, (y.shape[0], self.K), generator=self.gen).to(y.device)
            neg_idx = oc.remap_negatives_avid(raw, y)
        ve, ae = towers.av_forward(video, audio, self.sd, training=True)
        total, losses, self.Z = oc.criterion_forward(ve, ae, y, self.bank_v, self.bank_a, neg_idx, self.keys, self.Z)
        with torch.no_grad():
            oc.bank_update(self.bank_v, self.bank_a, ve, ae, y, self.momentum)
        loss = float(total.detach())
        self.opt.zero_grad()
        total.backward()
        self.opt.step()
        return loss
