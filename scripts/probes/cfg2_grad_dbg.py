import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_config2_gpu as T
from oracle import synth, towers
from avid_cma_b200 import models
from avid_cma_b200.criterions import AVID
from avid_cma_b200.models._tower import _MATH
DEV = "cuda:0"
seed = 22
B, N, K = T.B, T.N, T.K
video, audio = synth.clips(B, 8, 224, seed), synth.spectrograms(B, 200, 257, seed)
y = synth.instance_ids(B, N, seed); idx = synth.negatives(y, K, N, seed)
bank_v, bank_a = synth.bank(N, seed=seed, tag="bank_v"), synth.bank(N, seed=seed, tag="bank_a")
sd0 = synth.fill_state_dict(towers.state_dict_template(), seed=seed)
keys = [k for k in towers.param_keys(sd0)]
def oracle(dtype):
    sd = {k: (v.to(DEV, dtype) if v.is_floating_point() else v.to(DEV)) for k, v in sd0.items()}
    for k in keys: sd[k].requires_grad_(True)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    ve, ae = towers.av_forward(video.to(DEV, dtype), audio.to(DEV, dtype), sd, training=True)
    from oracle import criterion as oc
    total, losses, Z = oc.criterion_forward(ve, ae, y.to(DEV), bank_v.to(DEV, dtype), bank_a.to(DEV, dtype), idx.to(DEV), oc.avid_keys(K), -1.0)
    total.backward()
    return {k: sd[k].grad.detach().double().cpu() for k in keys}
g64 = oracle(torch.float64); g32 = oracle(torch.float32)
torch.cuda.empty_cache()
for math in ("bf16x3", "fp32"):
    model = models.av_wrapper('R2Plus1D', {'depth': 18}, 'Conv2D', {'depth': 10}, proj_dim=[512, 512, 128])
    model.load_state_dict(sd0)
    model.video_model.math = model.audio_model.math = _MATH[math]
    model = model.to(DEV).train()
    crit = AVID(num_data=N, embedding_dim=128, num_negatives=K, momentum=0.5, xModal_coeff=1., wModal_coeff=0., device=0)
    crit.nce_average.view1_mem.copy_(bank_v); crit.nce_average.view2_mem.copy_(bank_a)
    crit.nce_average.sample_negatives = lambda y_, K_: idx.to(DEV)
    ve, ae = model(video.to(DEV), audio.to(DEV))
    loss, log = crit(ve, ae, y.to(DEV)); loss.backward(); torch.cuda.synchronize()
    params = dict(model.named_parameters())
    print("==", math)
    rows = []
    for k in keys:
        a = params[k].grad.double().cpu(); w = g64[k]
        rows.append((float((a - w).norm() / w.norm().clamp_min(1e-30)), float((g32[k] - w).norm() / w.norm().clamp_min(1e-30)), k))
    for e, own, k in sorted(rows, reverse=True)[:14]: print("  %.2e (fp32 oracle %.1e)  %s" % (e, own, k))
    print("  median %.2e" % sorted(r[0] for r in rows)[len(rows) // 2])
    del model, crit; torch.cuda.empty_cache()
