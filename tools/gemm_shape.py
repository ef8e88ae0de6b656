"""One tower GEMM shape with its real epilogue: python tools/gemm_shape.py {qkv|out|fc1|fc2|tqkv|tout|tfc1|tfc2} (VLMCLIP_GEMM_DEBUG=1 prints phases)."""
import sys, torch
sys.path.insert(0, ".")
from vlm_clip_b200 import ops
dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "out"
shapes = {"qkv": (50432, 2304, 768, "fold"), "out": (50432, 768, 768, "res"), "fc1": (50432, 3072, 768, "fold_act"),
          "fc2": (50432, 768, 3072, "res"), "tqkv": (19712, 1536, 512, "fold"), "tout": (19712, 512, 512, "res"),
          "tfc1": (19712, 2048, 512, "fold_act"), "tfc2": (19712, 512, 2048, "res")}
M, N, K, kind = shapes[which]
a = torch.randn(M, K, device=dev).to(torch.bfloat16)
w = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
bias = torch.randn(N, device=dev); st = torch.rand(M, 2, device=dev); cc = torch.randn(N, device=dev)
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
res = torch.randn(M, N, device=dev).to(torch.bfloat16)
part = torch.empty(M, N // 32, 2, device=dev)
if kind == "fold_act": f = lambda: ops.gemm(a, w, bias=bias, row_stats=st, col_c=cc, act=1, out=out)
elif kind == "fold": f = lambda: ops.gemm(a, w, bias=bias, row_stats=st, col_c=cc, out=out)
else: f = lambda: ops.gemm(a, w, bias=bias, residual=res, out=res, stats_part_out=part)
for _ in range(2): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): f()
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 3
print(f"{which} {M}x{N}x{K}: {t*1e3:.1f} us  {2*M*N*K/t/1e9:.1f} TF/s")
