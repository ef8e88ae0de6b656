"""Sustained (power-capped) throughput of the tcgen05 GEMM vs cuBLAS on the tower shapes: ~1.5 s back to back each."""
import sys, time, torch
sys.path.insert(0, ".")
from vlm_clip_b200 import ops
dev = torch.device("cuda:0")
def sustained(fn, secs=1.5):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0; t0 = time.perf_counter(); e0.record()
    while time.perf_counter() - t0 < secs:
        for _ in range(20): fn()
        n += 20
        torch.cuda.synchronize() if n % 200 == 0 else None
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, M, N, K, kw in [("v.fc1 fold+qgelu", 50432, 3072, 768, "fold_act"), ("v.fc2 +res", 50432, 768, 3072, "res"),
                          ("v.qkv fold", 50432, 2304, 768, "fold"), ("v.out +res", 50432, 768, 768, "res"),
                          ("square 8192", 8192, 8192, 8192, "plain")]:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
    bias = torch.randn(N, device=dev); st = torch.rand(M, 2, device=dev); cc = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    res = torch.randn(M, N, device=dev).to(torch.bfloat16)
    if kw == "fold_act": f = lambda: ops.gemm(a, w, bias=bias, row_stats=st, col_c=cc, act=1, out=out)
    elif kw == "fold": f = lambda: ops.gemm(a, w, bias=bias, row_stats=st, col_c=cc, out=out)
    elif kw == "res": f = lambda: ops.gemm(a, w, bias=bias, residual=res, out=res)
    else: f = lambda: ops.gemm(a, w, out=out)
    wt = w.t()
    g = lambda: torch.matmul(a, wt, out=out)
    tm = sustained(f); tc = sustained(g)
    fl = 2.0 * M * N * K
    print(f"{name:18s} mine {tm*1e3:8.1f} us {fl/tm/1e9:7.1f} TF/s | cuBLAS plain {tc*1e3:8.1f} us {fl/tc/1e9:7.1f} TF/s", flush=True)
