"""CPU emulation of the bf16x3 operand split (x = hi + lo, products hi*hi + hi*lo + lo*hi, fp32 accumulate) through the
audio tower at BASELINE config 1, against fp64 -- the error budget behind the bf16x3 tolerances of
tests/test_towers_gpu.py.  Finding (see DESIGN.md "Precision"): embeddings move by ~1e-5, but a handful of ReLU gates
(4 of ~1M) flip sign, and each flip is an O(1) change of one gradient element, so early-layer gradients differ from fp64 by
~6e-3 in relative L2 -- independent of the dropped lo*lo term (mode x4 gives the same).  Not used by any test.

    python scripts/emulate_bf16x3_budget.py [x3|x4]
"""
import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
import torch.nn.functional as F
from oracle import synth, towers, criterion as oc
torch.set_num_threads(8)

def split(x, terms=2):
    hi = x.float().bfloat16().double()
    lo = (x - hi).float().bfloat16().double()
    return hi, lo

MODE = sys.argv[1] if len(sys.argv) > 1 else "x3"

class EmuConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, stride, padding):
        ctx.save_for_backward(x, w); ctx.stride, ctx.padding = stride, padding
        xh, xl = split(x); wh, wl = split(w)
        y = F.conv2d(xh, wh, stride=stride, padding=padding) + F.conv2d(xh, wl, stride=stride, padding=padding) + F.conv2d(xl, wh, stride=stride, padding=padding)
        if MODE == "x4": y = y + F.conv2d(xl, wl, stride=stride, padding=padding)
        return y.float().double()   # fp32 accumulator output
    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.float().double()
        dh, dl = split(dy); xh, xl = split(x); wh, wl = split(w)
        ci = torch.nn.grad.conv2d_input; cw = torch.nn.grad.conv2d_weight
        if ctx.stride == 1:
            dx = ci(x.shape, wh, dh, stride=1, padding=ctx.padding) + ci(x.shape, wl, dh, stride=1, padding=ctx.padding) + ci(x.shape, wh, dl, stride=1, padding=ctx.padding)
        else:
            dx = ci(x.shape, w.float().double(), dy, stride=ctx.stride, padding=ctx.padding)
        dw = cw(xh, w.shape, dh, stride=ctx.stride, padding=ctx.padding) + cw(xh, w.shape, dl, stride=ctx.stride, padding=ctx.padding) + cw(xl, w.shape, dh, stride=ctx.stride, padding=ctx.padding)
        return dx.float().double(), dw.float().double(), None, None

MASKS = []
def audio_tower(x, sd, emu):
    MASKS.append([])
    p = "audio_model"
    h = F.conv2d(x, sd[p + ".conv1.0.weight"], stride=2, padding=3)
    h = F.relu(towers._bn(h, sd, p + ".conv1.1", True))
    for (name, cin, cout, stride) in towers.AUDIO_BLOCKS:
        q = f"{p}.{name}"
        conv = (lambda a, w, s, pd: EmuConv.apply(a, w, s, pd)) if emu else (lambda a, w, s, pd: F.conv2d(a, w, stride=s, padding=pd))
        h = conv(h, sd[q + ".conv1.weight"], stride, 1)
        h = towers._bn(h, sd, q + ".bn1", True); MASKS[-1].append(h.detach() > 0); h = F.relu(h)
        h = conv(h, sd[q + ".conv2.weight"], 1, 1)
        h = towers._bn(h, sd, q + ".bn2", True); MASKS[-1].append(h.detach() > 0); h = F.relu(h)
    return F.adaptive_max_pool2d(h, 1)

def run(emu, dtype=torch.float64):
    sd = synth.fill_state_dict(towers.state_dict_template(), seed=0)
    sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    keys = [k for k in towers.param_keys(sd) if k.startswith("audio")]
    for k in keys: sd[k].requires_grad_(True)
    audio = synth.spectrograms(4, 100, 129, 0).to(dtype)
    a = audio_tower(audio, sd, emu)
    a = towers.head(a.view(4, 512), sd, "audio_proj")
    bank = synth.bank(64, seed=0, tag="bank_v").to(dtype)
    ah = F.normalize(a, dim=1)
    s = ah @ bank.t() / 0.07
    loss = -torch.log_softmax(s, 1)[torch.arange(4), torch.tensor([1, 17, 33, 60])].mean()
    loss.backward()
    return a.detach(), {k: sd[k].grad for k in keys}

e64, g64 = run(False)
e32, g32 = run(False, torch.float32)
ee, ge = run(True)
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
print("emb: fp32 err", rel(e32, e64), "emu err", rel(ee, e64))
for k in ["audio_model.conv1.0.weight", "audio_model.conv1.1.bias", "audio_model.conv1.1.weight", "audio_model.block1.conv1.weight", "audio_model.block2.conv2.weight", "audio_model.block4.conv2.weight"]:
    print(k, "fp32", rel(g32[k], g64[k]), MODE, rel(ge[k], g64[k]))

for i,(a,b,c) in enumerate(zip(MASKS[0], MASKS[1], MASKS[2])): print('layer', i, 'numel', a.numel(), 'flips fp32', int((a!=b).sum()), 'flips emu', int((a!=c).sum()))
