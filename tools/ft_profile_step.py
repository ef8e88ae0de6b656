"""Two full fine-tune steps (ViT-B/16, config 5) for ncu: tools/gpu_profile_ft.sh captures the backward kernels of the
second step.  usage: python tools/ft_profile_step.py [batch]"""
import sys
import torch
sys.path.insert(0, ".")
from vlm_clip_b200.model_m import CLIPWithAdapters
from vlm_clip_b200.trainer import CLIPAdapterTrainer
from vlm_clip_b200.configs import random_init_clip
from vlm_clip_b200.data import synthetic_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
clip = random_init_clip("openai/clip-vit-base-patch16", seed=0).to(dev)
model = CLIPWithAdapters(clip=clip, freeze_clip=False, use_text_adapter=False, use_vision_adapter=False,
                         use_shared_adapters=False).to(dev)
model.train()
pix, ids, mask = synthetic_batch(B, seed=2)
batch = {"input_ids": ids.to(dev), "attention_mask": mask.to(dev), "pixel_values": pix.to(dev)}
tr = CLIPAdapterTrainer(model, [batch], learning_rate=1e-7, output_dir="/tmp/vlmclip_ft_prof", trainable="all")
for _ in range(2):
    loss = tr.training_step(batch)
torch.cuda.synchronize()
print("loss", loss.item())
