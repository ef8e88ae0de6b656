import sys, torch
sys.path.insert(0, ".")
from vlm_clip_b200 import ops
dev = torch.device("cuda:0")
M, N, K = 50432, 768, 3072
a = torch.randn(M, K, device=dev).to(torch.bfloat16)
w = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    ops.gemm(a, w, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.gemm(a, w, out=out)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 5
print(f"gemm {M}x{N}x{K}: {t*1e3:.1f} us  {2*M*N*K/t/1e9:.1f} TF/s")
