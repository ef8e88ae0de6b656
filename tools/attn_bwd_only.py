"""Times ops.attention_bwd (the two mma.sync backward kernels) and the mma.sync forward on the full-fine-tune shapes:
python tools/attn_bwd_only.py  (vision B=256 S=197 H=12 unmasked; text B=256 S=77 H=8 causal)."""
import sys, torch
sys.path.insert(0, ".")
from vlm_clip_b200 import ops
dev = torch.device("cuda:0")


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for (B, S, H, causal) in [(256, 197, 12, False), (256, 77, 8, True), (64, 257, 16, False)]:
    qkv = torch.randn(B * S, 3 * H * 64, device=dev).to(torch.bfloat16)
    dout = torch.randn(B * S, H * 64, device=dev).to(torch.bfloat16)
    out = ops.attention(qkv, B, S, H, causal=causal)
    tf = timeit(lambda: ops.attention(qkv, B, S, H, causal=causal, out=out))
    tb = timeit(lambda: ops.attention_bwd(qkv, out, dout, B, S, H, causal=causal))
    mm = 2.0 * B * H * S * S * 64
    print(f"B={B} S={S} H={H} causal={causal}: fwd {tf:7.1f} us ({2 * mm / tf * 1e-6:6.1f} TF/s)   "
          f"bwd {tb:7.1f} us ({8 * mm / tb * 1e-6:6.1f} TF/s of 8 matmul-equivalents)", flush=True)
