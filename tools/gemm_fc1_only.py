import sys, torch
sys.path.insert(0, ".")
from vlm_clip_b200 import ops
dev = torch.device("cuda:0")
M, N, K = 50432, 3072, 768
a = torch.randn(M, K, device=dev).to(torch.bfloat16)
w = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
bias = torch.randn(N, device=dev); st = torch.rand(M, 2, device=dev); cc = torch.randn(N, device=dev)
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
mode = sys.argv[1] if len(sys.argv) > 1 else "full"
kw = dict(bias=bias, row_stats=st, col_c=cc, act=1) if mode == "full" else (dict(bias=bias, act=1) if mode == "act" else (dict(bias=bias, row_stats=st, col_c=cc) if mode == "fold" else {}))
for _ in range(2):
    ops.gemm(a, w, out=out, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    ops.gemm(a, w, out=out, **kw)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 3
print(f"fc1 mode={mode}: {t*1e3:.1f} us  {2*M*N*K/t/1e9:.1f} TF/s")
