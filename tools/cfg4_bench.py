"""BASELINE config 4: video clips, ViT-B/16 + bottleneck adapters, 64 clips x 8 frames (+ 64 captions) per step, temporal
mean-pool of the per-frame features, Track-M train step.  Times the float path ([B, 3, T, H, W], stacked `process_video`
outputs) and the uint8 path ([B, T, Hs, Ws, 3] decoded frames, preprocessing fused into the patch extraction).
Development tool, not the judged bench.   usage: python tools/cfg4_bench.py [clips] [frames] [steps]"""
import sys, json
import torch
sys.path.insert(0, ".")
from vlm_clip_b200.model_m import CLIPWithAdapters
from vlm_clip_b200.trainer import CLIPAdapterTrainer
from vlm_clip_b200.configs import flops_per_pair, random_init_clip
from vlm_clip_b200.data import synthetic_batch

B16 = "openai/clip-vit-base-patch16"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
dev = torch.device("cuda:0")
clip = random_init_clip(B16, seed=0).to(dev)
for p in clip.parameters():
    p.requires_grad_(False)
torch.manual_seed(1)
model = CLIPWithAdapters(clip=clip, use_shared_adapters=False).to(dev)
model.train()
g = torch.Generator().manual_seed(4)
ids = torch.randint(3, 49406, (B, 77), generator=g)
ids[:, 0], ids[:, -1] = 49406, 49407
common = {"input_ids": ids.to(dev), "attention_mask": torch.ones(B, 77, dtype=torch.int64, device=dev)}
clips_f = torch.randn(B, 3, T, 224, 224, generator=g).to(dev)
clips_u8 = torch.randint(0, 256, (B, T, 360, 480, 3), generator=g, dtype=torch.uint8).to(dev)
tr = CLIPAdapterTrainer(model, [None], output_dir="/tmp/vlmclip_cfg4")
fl = flops_per_pair(B16)
res = {"workload": "config 4: ViT-B/16 + adapters, video clips with temporal mean-pool, Track-M train step", "clips": B,
       "frames_per_clip": T}
for tag, pix in (("float_clips", clips_f), ("uint8_frames_360x480", clips_u8)):
    batch = dict(common, pixel_values=pix)
    for _ in range(3):
        tr.training_step(batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = tr.training_step(batch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    tf = (fl["image"] * B * T + fl["caption"] * B) / 1e12
    res[tag] = {"ms_per_step": ms, "clips_per_s": B / ms * 1e3, "frames_per_s": B * T / ms * 1e3,
                "step_tflops": tf / (ms / 1e3), "frac_of_sustained_peak_1356.7": tf / (ms / 1e3) / 1356.7, "loss": loss.item()}
print(json.dumps(res))
