"""BASELINE config 3 on ONE GPU's share: CLIP ViT-L/14 + PE-CLIP adapters, 512 pairs per GPU (global batch 4096 on 8 GPUs).
Times (a) the per-GPU train step with the local 512 x 512 loss and (b) the global 4096 x 4096 loss + local-row gradient
kernel on its own (what every rank runs after the all-gather), so the 8-GPU step is (a) - loss(512) + (b) + two small
NCCL calls.  Development tool, not the judged bench.   usage: python tools/cfg3_bench.py [batch_per_gpu] [steps]"""
import sys, json
import torch
sys.path.insert(0, ".")
from vlm_clip_b200 import ops
from vlm_clip_b200.model_m import CLIPWithAdapters
from vlm_clip_b200.trainer import CLIPAdapterTrainer
from vlm_clip_b200.configs import flops_per_pair, random_init_clip
from vlm_clip_b200.data import synthetic_batch

L14 = "openai/clip-vit-large-patch14"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
clip = random_init_clip(L14, seed=0).to(dev)
for p in clip.parameters():
    p.requires_grad_(False)
torch.manual_seed(1)
model = CLIPWithAdapters(clip=clip, use_shared_adapters=False, adapter_kind="peclip").to(dev)
model.train()
g = torch.Generator().manual_seed(3)
batches = []
for _ in range(2):
    ids = torch.randint(3, 49406, (B, 77), generator=g)
    ids[:, 0], ids[:, -1] = 49406, 49407
    batches.append({"input_ids": ids.to(dev), "attention_mask": torch.ones(B, 77, dtype=torch.int64, device=dev),
                    "pixel_values": torch.randn(B, 3, 224, 224, generator=g).to(dev).to(torch.bfloat16)})
tr = CLIPAdapterTrainer(model, [None], output_dir="/tmp/vlmclip_cfg3")
for i in range(3):
    tr.training_step(batches[i % 2])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    loss = tr.training_step(batches[i % 2])
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
fl = flops_per_pair(L14)
res = {"workload": "config 3 per-GPU share: ViT-L/14 + PE-CLIP adapters, Track-M train step", "batch_per_gpu": B,
       "ms_per_step": ms, "images_per_s_per_gpu": B / ms * 1e3, "step_tflops": fl["pair"] * B / ms / 1e9,
       "frac_of_sustained_peak_1356.7": fl["pair"] * B / ms / 1e9 / 1356.7, "loss": loss.item(),
       "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
# global loss: N = 8 * B rows gathered, this rank differentiates its own B rows
Ng, P = 8 * B, 768
t = torch.randn(Ng, P, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
im = torch.randn(Ng, P, device=dev, generator=torch.Generator(device=dev).manual_seed(6))
tl, il = t[:B].clone().requires_grad_(True), im[:B].clone().requires_grad_(True)
for _ in range(2):
    out = ops.clip_loss(tl, il, 100.0, t, im, 0, want_logits=False)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    out = ops.clip_loss(tl, il, 100.0, t, im, 0, want_logits=False)
e1.record()
torch.cuda.synchronize()
res["global_loss_4096_ms"] = e0.elapsed_time(e1) / 5
res["global_loss_value"] = out[0].item()
print(json.dumps(res))
ops.TRACE = []
tr.training_step(batches[0])
torch.cuda.synchronize()
import collections
agg = collections.defaultdict(lambda: [0.0, 0])
for name, a, b, extra in ops.TRACE:
    agg[name][0] += a.elapsed_time(b)
    agg[name][1] += 1
ops.TRACE = None
for k, (tt, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:8]:
    print(f"{k:24s} {tt:9.2f} ms  {n:5d} calls")
