import sys, torch
sys.path.insert(0, ".")
from vlm_clip_b200 import ops
dev = torch.device("cuda:0")
B, S, H = 256, 197, 12
qkv = torch.randn(B * S, 3 * H * 64, device=dev).to(torch.bfloat16)
out = torch.empty(B * S, H * 64, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(qkv, B, S, H, out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    ops.attention(qkv, B, S, H, out=out)
b.record(); torch.cuda.synchronize()
print("attention vision us", a.elapsed_time(b) / 5 * 1e3)
