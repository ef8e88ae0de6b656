"""Times ops.attention alone: python tools/attn_only.py [B S H] (default: the ViT-B/16 vision tower of the headline).
VLMCLIP_ATTN_SPLIT=0 keeps 224 < S <= 384 on the mma.sync kernel (A/B against the tcgen05 key-range split)."""
import sys, torch
sys.path.insert(0, ".")
from vlm_clip_b200 import ops
dev = torch.device("cuda:0")
B, S, H = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (256, 197, 12)
qkv = torch.randn(B * S, 3 * H * 64, device=dev).to(torch.bfloat16)
out = torch.empty(B * S, H * 64, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(qkv, B, S, H, out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    ops.attention(qkv, B, S, H, out=out)
b.record(); torch.cuda.synchronize()
us = a.elapsed_time(b) / 10 * 1e3
flops = 4.0 * B * H * S * S * 64
print(f"attention B={B} S={S} H={H}: {us:.1f} us  ({flops / us * 1e-6:.1f} TFLOP/s)")
