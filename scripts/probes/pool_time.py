import sys, torch
sys.path.insert(0, "/root/repo")
from avid_cma_b200 import ops
dev = "cuda"
n, t, h, w, c = 64, 8, 112, 112, 64
z = torch.randn(n, t, h, w, c, device=dev)
gamma, beta = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev) * 0.1
rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
st = ops.bn_train_stats(z, gamma, beta, rm, rv)

def timeit(name, fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    print("%-40s %.3f ms" % (name, e0.elapsed_time(e1) / reps))

y = ops.bn_relu_forward_act(z, st.scale, st.shift, True, False, False)
p, am = ops.maxpool_1x3x3_forward(y.f32)
dyp = torch.randn_like(p)
timeit("bn_relu_forward (f32 out)", lambda: ops.bn_relu_forward_act(z, st.scale, st.shift, True, False, False))
timeit("maxpool fwd", lambda: ops.maxpool_1x3x3_forward(y.f32))
timeit("split planes of pooled", lambda: ops.split_bf16(p))
timeit("FUSED bn_relu_maxpool fwd (f32+planes)", lambda: ops.bn_relu_maxpool_forward(z, st.scale, st.shift, True, True, True))
dy = ops.maxpool_1x3x3_backward(am, dyp, y.f32.shape)
timeit("maxpool bwd", lambda: ops.maxpool_1x3x3_backward(am, dyp, y.f32.shape))
timeit("bn bwd reduce+apply (planes)", lambda: ops.bn_relu_backward_act(z, dy, st, gamma, beta, False, True, True))
timeit("FUSED pool bn bwd reduce+apply", lambda: ops.bn_relu_maxpool_backward_act(z, p, am, dyp, st, gamma, beta, False, True, True))
