"""Residual GEMM (out-proj / fc2) with the two-term hi + lo stream vs the one-plane bf16 stream, ViT-B/16 @ B = 256 and
ViT-L/14 @ B = 512 shapes; CUDA events, L2 flushed between iterations.  VLMCLIP_GEMM_RES2_CFG selects the pair-kernel
pipeline (52 default / 43 / 61); run once per value.  Development tool."""
import os
import sys

import torch

sys.path.insert(0, ".")
from vlm_clip_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
bf16 = torch.bfloat16
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=12, warm=3):
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


def case(name, M, N, K):
    a = torch.randn(M, K, device=dev).to(bf16)
    w = (torch.randn(N, K, device=dev) * 0.02).to(bf16)
    bias = torch.randn(N, device=dev)
    x2 = torch.randn(2, M, N, device=dev).to(bf16)
    x1 = torch.randn(M, N, device=dev).to(bf16)
    part = torch.empty(M, N // 32, 2, device=dev)
    t2 = timeit(lambda: ops.gemm_res2(a, w, bias, x2, stats_part_out=part))
    t1 = timeit(lambda: ops.gemm(a, w, bias=bias, residual=x1, out=x1, stats_part_out=part))
    fl = 2.0 * M * N * K
    by2 = (M * K + N * K) * 2 + M * N * 8 + M * N // 32 * 8
    print(f"{name:14s} M={M:6d} N={N:5d} K={K:5d} cfg={os.environ.get('VLMCLIP_GEMM_RES2_CFG', '52')}: hi+lo {t2 * 1e3:7.1f} us "
          f"{fl / t2 / 1e9:7.1f} TF/s {by2 / t2 / 1e6:6.0f} GB/s | bf16 stream {t1 * 1e3:7.1f} us {fl / t1 / 1e9:7.1f} TF/s", flush=True)


case("B16 out-proj", 50432, 768, 768)
case("B16 fc2", 50432, 768, 3072)
case("B16 t.out", 19712, 512, 512)
case("B16 t.fc2", 19712, 512, 2048)
case("L14 out-proj", 131584, 1024, 1024)
case("L14 fc2", 131584, 1024, 4096)
