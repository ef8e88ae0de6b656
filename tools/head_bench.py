"""Isolated timings of the trainable-path (fp32) ops at the headline batch (256 rows): adapter fwd/bwd, projections,
contrastive loss fwd+bwd, fused AdamW.  CUDA events, 20 iterations each after warm-up."""
import sys, torch
sys.path.insert(0, ".")
from vlm_clip_b200 import ops, _native as N
dev = torch.device("cuda:0")
B = 256
def timeit(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e3
g = torch.Generator(device=dev).manual_seed(0)
def r(*s, scale=1.0, grad=False):
    t = torch.randn(*s, device=dev, generator=g) * scale
    return t.requires_grad_(grad)
for name, D in (("vision", 768), ("text", 512)):
    A = 256
    x = r(B, D)
    W1, b1, W2, b2 = r(A, D, scale=0.03, grad=True), r(A, grad=True), r(D, A, scale=0.05, grad=True), r(D, grad=True)
    gm, bt = torch.ones(D, device=dev, requires_grad=True), torch.zeros(D, device=dev, requires_grad=True)
    fwd = lambda: ops.adapter(x, W1, b1, W2, b2, gm, bt, act=N.ACT_GELU_ERF, post=N.POST_RESIDUAL_LN)
    t_f = timeit(fwd)
    y = fwd(); dy = torch.randn_like(y)
    def fb():
        yy = fwd(); yy.backward(dy)
    t_fb = timeit(fb)
    print(f"adapter {name:6s} D={D}: fwd {t_f:7.1f} us   fwd+bwd {t_fb:7.1f} us")
    P = 512
    Wp = r(P, D, scale=0.03)
    xg = r(B, D, grad=True)
    t_l = timeit(lambda: ops.linear_f32(xg, Wp))
    o = ops.linear_f32(xg, Wp); do = torch.randn_like(o)
    def lfb():
        oo = ops.linear_f32(xg, Wp); oo.backward(do)
    print(f"linear_f32 {D}->{P}: fwd {t_l:7.1f} us   fwd+bwd {timeit(lfb):7.1f} us")
t, i = r(B, 512, grad=True), r(B, 512, grad=True)
def lossfb():
    l = ops.clip_loss(t, i, 100.0, None, None, 0)[0]; l.backward()
print(f"clip_loss fwd     {timeit(lambda: ops.clip_loss(t, i, 100.0, None, None, 0)):7.1f} us   fwd+bwd {timeit(lossfb):7.1f} us")
params = [torch.nn.Parameter(r(256, 768)), torch.nn.Parameter(r(768, 256)), torch.nn.Parameter(r(256, 512)), torch.nn.Parameter(r(512, 256))]
opt = ops.FusedAdamW(params, lr=1e-4, max_grad_norm=1.0)
opt.grad.normal_()
print(f"adamw+clip step   {timeit(opt.step):7.1f} us  ({sum(p.numel() for p in params)} params)")
