"""Host-side time of each phase of training_step (no synchronisation added): finds calls that block on the GPU."""
import sys, time, torch
sys.path.insert(0, ".")
from vlm_clip_b200.configs import random_init_clip
from vlm_clip_b200.data import synthetic_batch
from vlm_clip_b200.model_m import CLIPWithAdapters
from vlm_clip_b200.trainer import CLIPAdapterTrainer
from vlm_clip_b200.dist import allreduce_sum_
dev = torch.device("cuda:0")
B = 256
clip = random_init_clip("openai/clip-vit-base-patch16", seed=0).to(dev)
model = CLIPWithAdapters(clip=clip, use_shared_adapters=False).to(dev)
trainer = CLIPAdapterTrainer(model, train_dataloader=[None], output_dir="/tmp/ab_ckpt")
pix, ids, mask = synthetic_batch(B)
batch = {"pixel_values": pix.to(dev), "input_ids": ids.to(dev), "attention_mask": mask.to(dev)}
torch.cuda.synchronize()
ready = torch.cuda.Event(); ready.record()
batch["inputs_ready"] = ready
for _ in range(5):
    trainer.training_step(batch)
torch.cuda.synchronize()
acc = {}
def tick(name, t0):
    t = time.perf_counter(); acc[name] = acc.get(name, 0.0) + (t - t0); return t
n = 30
for _ in range(n):
    t = time.perf_counter()
    out = model(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"], pixel_values=batch["pixel_values"],
                return_loss=True, inputs_ready=ready)
    t = tick("forward", t)
    loss = out["loss"]
    trainer.optimizer.zero_grad(); t = tick("zero_grad", t)
    loss.backward(); t = tick("backward", t)
    allreduce_sum_(trainer.optimizer.grad); t = tick("allreduce", t)
    trainer.optimizer.step(); t = tick("opt.step", t)
torch.cuda.synchronize()
for k, v in acc.items():
    print(f"{k:10s} {v / n * 1e3:7.3f} ms/step")
# forward split
import vlm_clip_b200.ops as ops
acc.clear()
bb = model._backbone()
for _ in range(n):
    t = time.perf_counter()
    th = bb.text_hidden_pre_ln(batch["input_ids"], batch["attention_mask"]); t = tick("text tower", t)
    vh = bb.vision_hidden(batch["pixel_values"]); t = tick("vision tower", t)
    tf = model._text_head(bb, th, B, 77); t = tick("text head", t)
    vf = model._image_head(bb, vh, B); t = tick("image head", t)
    l = ops.clip_loss(tf, vf, 100.0, None, None, 0); t = tick("loss", t)
torch.cuda.synchronize()
for k, v in acc.items():
    print(f"{k:12s} {v / n * 1e3:7.3f} ms/step")
