"""A/B of the stream schedules inside one process (interleaved rounds cancel the power-cap drift between boxes)."""
import sys, time, torch
sys.path.insert(0, ".")
from vlm_clip_b200.configs import random_init_clip
from vlm_clip_b200.data import synthetic_batch
from vlm_clip_b200.model_m import CLIPWithAdapters
from vlm_clip_b200.trainer import CLIPAdapterTrainer
dev = torch.device("cuda:0")
B = 256
clip = random_init_clip("openai/clip-vit-base-patch16", seed=0).to(dev)
model = CLIPWithAdapters(clip=clip, use_shared_adapters=False).to(dev)
trainer = CLIPAdapterTrainer(model, train_dataloader=[None], output_dir="/tmp/ab_ckpt")
pix, ids, mask = synthetic_batch(B)
batch = {"pixel_values": pix.to(dev), "input_ids": ids.to(dev), "attention_mask": mask.to(dev)}
torch.cuda.synchronize()
ready = torch.cuda.Event(); ready.record()
def run(mode, n):
    model.overlap_towers = mode != "serial"
    b = dict(batch)
    if mode == "pipelined":
        b["inputs_ready"] = ready
    for _ in range(3):
        trainer.training_step(b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        trainer.training_step(b)
    e1.record()
    t_cpu = time.perf_counter() - t0
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, t_cpu / n * 1e3
modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ("serial", "two_streams", "pipelined")
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
import os
head_stream = torch.cuda.Stream(priority=-1) if os.environ.get("HEAD_HIGH_PRIORITY") == "1" else torch.cuda.current_stream()
for r in range(rounds):
    for mode in modes:
        with torch.cuda.stream(head_stream):
            g, c = run(mode, 60)
        print(f"round {r} {mode:12s} gpu {g:7.3f} ms/step   cpu enqueue {c:6.2f} ms/step", flush=True)
