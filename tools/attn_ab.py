"""A/B of attention variants inside one process: isolated kernel and whole pipelined step (alternating rounds)."""
import os, sys, torch
sys.path.insert(0, ".")
from vlm_clip_b200 import ops
from vlm_clip_b200.configs import random_init_clip
from vlm_clip_b200.data import synthetic_batch
from vlm_clip_b200.model_m import CLIPWithAdapters
from vlm_clip_b200.trainer import CLIPAdapterTrainer
dev = torch.device("cuda:0")
B, S, H = 256, 197, 12
qkv = torch.randn(B * S, 3 * H * 64, device=dev).to(torch.bfloat16)
out = torch.empty(B * S, H * 64, device=dev, dtype=torch.bfloat16)
def t_attn(n=20):
    for _ in range(3): ops.attention(qkv, B, S, H, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): ops.attention(qkv, B, S, H, out=out)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
clip = random_init_clip("openai/clip-vit-base-patch16", seed=0).to(dev)
model = CLIPWithAdapters(clip=clip, use_shared_adapters=False).to(dev)
trainer = CLIPAdapterTrainer(model, train_dataloader=[None], output_dir="/tmp/ab_ckpt")
pix, ids, mask = synthetic_batch(B)
batch = {"pixel_values": pix.to(dev), "input_ids": ids.to(dev), "attention_mask": mask.to(dev)}
torch.cuda.synchronize()
ready = torch.cuda.Event(); ready.record(); batch["inputs_ready"] = ready
def t_step(n=60):
    for _ in range(3): trainer.training_step(batch)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): trainer.training_step(batch)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
var = sys.argv[1] if len(sys.argv) > 1 else "VLMCLIP_ATTN_TWO_PASS"
for r in range(3):
    for v in ("0", "1"):
        os.environ[var] = v
        print(f"round {r} {var}={v}: attention {t_attn():7.1f} us   step {t_step():7.3f} ms", flush=True)
