"""Per-kernel timing on a B200 (CUDA events, L2 flushed between iterations) for the ViT-B/16 @ B=256 shapes.
Prints achieved TFLOP/s or GB/s next to cuBLAS / torch for context.  Development tool, not the judged bench."""
import sys, time, json
import torch
sys.path.insert(0, ".")
from vlm_clip_b200 import ops, _native as N
from vlm_clip_b200.configs import random_init_clip
from vlm_clip_b200.data import synthetic_batch

dev = torch.device("cuda:0")
bf16 = torch.bfloat16
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def gemm_case(name, M, Nn, K, act=0, res=False, stats=False):
    a = torch.randn(M, K, device=dev).to(bf16)
    w = (torch.randn(Nn, K, device=dev) * 0.02).to(bf16)
    bias = torch.randn(Nn, device=dev)
    r = torch.randn(M, Nn, device=dev).to(bf16) if res else None
    st = torch.rand(M, 2, device=dev) if stats else None
    cc = torch.randn(Nn, device=dev) if stats else None
    out = torch.empty(M, Nn, device=dev, dtype=bf16)
    t = timeit(lambda: ops.gemm(a, w, bias=bias, residual=r, act=act, row_stats=st, col_c=cc, out=out))
    t0 = timeit(lambda: ops.gemm(a, w, out=out))
    tc = timeit(lambda: torch.matmul(a, w.t(), out=out))
    fl = 2.0 * M * Nn * K
    print(f"{name:28s} M={M:6d} N={Nn:5d} K={K:5d}  mine {t*1e3:8.1f} us {fl/t/1e9:7.1f} TF/s | plain {t0*1e3:8.1f} us {fl/t0/1e9:7.1f} | cuBLAS {tc*1e3:8.1f} us {fl/tc/1e9:7.1f}", flush=True)
    return t


B, S, D, F, H = 256, 197, 768, 3072, 12
M = B * S
tot = 0
tot += gemm_case("v.qkv (LN fold)", M, 3 * D, D, stats=True)
tot += gemm_case("v.out (+res)", M, D, D, res=True)
tot += gemm_case("v.fc1 (fold+qgelu)", M, F, D, act=1, stats=True)
tot += gemm_case("v.fc2 (+res)", M, D, F, res=True)
Mt, Dt, Ft = B * 77, 512, 2048
tt = 0
tt += gemm_case("t.qkv", Mt, 3 * Dt, Dt, stats=True)
tt += gemm_case("t.out", Mt, Dt, Dt, res=True)
tt += gemm_case("t.fc1", Mt, Ft, Dt, act=1, stats=True)
tt += gemm_case("t.fc2", Mt, Dt, Ft, res=True)
gemm_case("patch", B * 196, D, 768)
print(f"vision GEMMs/layer {tot:.3f} ms -> x12 = {tot*12:.2f} ms ; text {tt:.3f} ms -> x12 = {tt*12:.2f} ms")

qkv = torch.randn(M, 3 * D, device=dev).to(bf16)
att = torch.empty(M, D, device=dev, dtype=bf16)
t = timeit(lambda: ops.attention(qkv, B, S, H, out=att))
print(f"attention vision  {t*1e3:8.1f} us  {4.0*B*H*S*S*64/t/1e9:7.1f} TF/s")
q4 = qkv.view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4).contiguous()
t2 = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q4[0], q4[1], q4[2]))
print(f"   torch SDPA     {t2*1e3:8.1f} us")
qkvt = torch.randn(Mt, 3 * Dt, device=dev).to(bf16)
attt = torch.empty(Mt, Dt, device=dev, dtype=bf16)
t3 = timeit(lambda: ops.attention(qkvt, B, 77, 8, causal=True, out=attt))
print(f"attention text    {t3*1e3:8.1f} us")
x = torch.randn(M, D, device=dev).to(bf16)
st = torch.empty(M, 2, device=dev)
t4 = timeit(lambda: ops.row_stats(x, out=st))
print(f"row_stats vision  {t4*1e3:8.1f} us  {M*D*2/t4/1e6:7.1f} GB/s")
g = torch.ones(D, device=dev); b = torch.zeros(D, device=dev); y = torch.empty_like(x)
t5 = timeit(lambda: ops.layernorm(x, g, b, out=y))
print(f"layernorm vision  {t5*1e3:8.1f} us  {M*D*4/t5/1e6:7.1f} GB/s")
pix = torch.randn(B, 3, 224, 224, device=dev)
t6 = timeit(lambda: ops.im2col(pix, 16))
print(f"im2col            {t6*1e3:8.1f} us  {(pix.numel()*4+B*196*768*2)/t6/1e6:7.1f} GB/s")

# whole model step
from vlm_clip_b200.model_m import CLIPWithAdapters
from vlm_clip_b200.trainer import CLIPAdapterTrainer
clip = random_init_clip("openai/clip-vit-base-patch16", seed=0).to(dev)
torch.manual_seed(1)
model = CLIPWithAdapters(clip=clip, use_shared_adapters=False).to(dev)
model.train()
pixs, ids, mask = synthetic_batch(B)
batch = {"input_ids": ids.to(dev), "attention_mask": mask.to(dev), "pixel_values": pixs.to(dev)}
tr = CLIPAdapterTrainer(model, [None], output_dir="/tmp/vlmclip_kb")
n0 = N.launch_count()
ts = timeit(lambda: tr.training_step(batch), iters=8, warm=3)
n1 = N.launch_count()
print(f"train step B=256 ViT-B/16: {ts:.2f} ms -> {B/ts*1e3:.0f} img/s ; launches/step ~{(n1-n0)/11:.0f}")
bb = model._backbone()
tv = timeit(lambda: bb.vision_hidden(batch['pixel_values']), iters=8)
ttx = timeit(lambda: bb.text_hidden_pre_ln(batch['input_ids'], batch['attention_mask']), iters=8)
print(f"vision tower {tv:.2f} ms ; text tower {ttx:.2f} ms")
# per-op trace of one vision tower forward
ops.TRACE = []
bb.vision_hidden(batch['pixel_values'])
torch.cuda.synchronize()
import collections
agg = collections.OrderedDict()
for name, e0, e1, extra in ops.TRACE:
    key = f"{name} {extra}"
    agg.setdefault(key, []).append(e0.elapsed_time(e1) * 1e3)
ops.TRACE = None
tot = 0
for key, v in agg.items():
    print(f"   {key:48s} n={len(v):3d} avg {sum(v)/len(v):8.1f} us  total {sum(v)/1e3:7.3f} ms")
    tot += sum(v)
print(f"   traced total {tot/1e3:.2f} ms")
# autocast comparator: error of torch's own bf16 path against fp32 on the same weights
with torch.no_grad():
    ps, idss, ms = batch['pixel_values'][:8], batch['input_ids'][:8], batch['attention_mask'][:8]
    ref = clip.vision_model(pixel_values=ps).last_hidden_state
    with torch.autocast("cuda", dtype=bf16):
        ac = clip.vision_model(pixel_values=ps).last_hidden_state
    mine = bb.vision_hidden(ps).view(8, S, D)
    rel = lambda a, b: ((a.float()-b.float()).norm()/b.float().norm()).item()
    print(f"vision last_hidden rel err vs fp32: autocast {rel(ac, ref):.2e}  native {rel(mine, ref):.2e}")
    hb = clip.half().bfloat16() if False else None
