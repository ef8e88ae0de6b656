"""Full fine-tune step (BASELINE config 5: ViT-B/16, adapters disabled, all CLIP parameters trainable) on one B200:
step time, peak memory and the per-op breakdown (ops.TRACE, CUDA events).  Development tool, not the judged bench.
usage: python tools/ft_bench.py [batch] [steps]"""
import sys, json, collections
import torch
sys.path.insert(0, ".")
from vlm_clip_b200 import ops
from vlm_clip_b200.model_m import CLIPWithAdapters
from vlm_clip_b200.trainer import CLIPAdapterTrainer
from vlm_clip_b200.configs import flops_per_pair, random_init_clip
from vlm_clip_b200.data import synthetic_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
clip = random_init_clip("openai/clip-vit-base-patch16", seed=0).to(dev)
model = CLIPWithAdapters(clip=clip, freeze_clip=False, use_text_adapter=False, use_vision_adapter=False,
                         use_shared_adapters=False).to(dev)
model.train()
pix, ids, mask = synthetic_batch(B, seed=2)
ids[:, 0] = torch.arange(B) % 1000 + 5
batch = {"input_ids": ids.to(dev), "attention_mask": mask.to(dev), "pixel_values": pix.to(dev).to(torch.bfloat16)}
tr = CLIPAdapterTrainer(model, [batch], learning_rate=1e-7, output_dir="/tmp/vlmclip_ft_bench", trainable="all")
for _ in range(2):
    loss = tr.training_step(batch)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = tr.training_step(batch)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
fl = flops_per_pair("openai/clip-vit-base-patch16")
res = {"workload": "CLIP ViT-B/16 full fine-tune (adapters disabled), Track-M step", "batch": B, "ms_per_step": ms,
       "images_per_s": B / ms * 1e3, "loss": loss.item(), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}
try:
    res["step_tflops"] = 3 * fl["pair"] * B / ms / 1e9
except Exception:
    pass
print(json.dumps(res))
ops.TRACE = []
tr.training_step(batch)
torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for name, a, b, extra in ops.TRACE:
    key = name if name != "gemm" else ("gemm_wgrad" if "->" in extra and int(extra.split("x")[1].split("-")[0]) > 10000 else "gemm")
    agg[key][0] += a.elapsed_time(b)
    agg[key][1] += 1
ops.TRACE = None
for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:24s} {t:9.2f} ms  {n:5d} calls")
