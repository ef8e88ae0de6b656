"""Interleaved A/B of one switch inside ONE process (the power-capped clock drifts by a few percent between processes
and between legs of a run, which hides 1-2 % effects): rounds of N eager train steps, alternating the variants.

    python tools/ab_step.py fused_stats|residual|pdl_tail [steps] [rounds]

fused_stats  residual GEMMs finalise the LayerNorm statistics vs separate ln_partials_to_stats launches
residual     two-term (hi + lo) residual stream vs one bf16 plane
graph        two-graph replay vs eager launches
Development tool, not the judged bench."""
import statistics
import sys

import torch

sys.path.insert(0, ".")
from vlm_clip_b200 import towers  # noqa: E402
from vlm_clip_b200.configs import random_init_clip  # noqa: E402
from vlm_clip_b200.model_m import CLIPWithAdapters  # noqa: E402
from vlm_clip_b200.trainer import CLIPAdapterTrainer  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "fused_stats"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dev = torch.device("cuda:0")
clip = random_init_clip("openai/clip-vit-base-patch16", seed=0).to(dev)
B = 256
g = torch.Generator().manual_seed(1)
batches = []
for _ in range(3):
    ids = torch.randint(3, 49406, (B, 77), generator=g)
    ids[:, 0], ids[:, -1] = 49406, 49407
    batches.append({"input_ids": ids.to(dev), "attention_mask": torch.ones(B, 77, dtype=torch.int64, device=dev),
                    "pixel_values": torch.randn(B, 3, 224, 224, generator=g).to(dev)})
torch.cuda.synchronize()
ready = torch.cuda.Event()
ready.record()
for b in batches:
    b["inputs_ready"] = ready


def make(residual=None, graph=False):
    torch.manual_seed(1)
    m = CLIPWithAdapters(clip=clip, use_shared_adapters=False).to(dev).train()
    if residual is not None:
        from vlm_clip_b200.towers import NativeClipTowers

        m._backbone()  # key the cache, then swap the towers object
        m._towers = NativeClipTowers(clip, dev, residual=residual)
    return CLIPAdapterTrainer(m, [None], output_dir="/tmp/vlmclip_ab", cuda_graph=graph)


if what == "fused_stats":
    tr = make()
    variants = {"fused": lambda: setattr(towers, "_FUSED_STATS", True), "separate": lambda: setattr(towers, "_FUSED_STATS", False)}
    trainers = {"fused": tr, "separate": tr}
elif what == "residual":
    trainers = {"hilo": make("hilo"), "bf16": make("bf16")}
    variants = {k: (lambda: None) for k in trainers}
elif what == "graph":
    trainers = {"graph": make(graph=True), "eager": make(graph=False)}
    variants = {k: (lambda: None) for k in trainers}
else:
    raise SystemExit(__doc__)


def run(name, n):
    variants[name]()
    tr = trainers[name]
    for i in range(6):
        tr.training_step(batches[i % 3])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        tr.training_step(batches[i % 3])
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for name in trainers:  # settle: power-capped steady state, graphs captured
    run(name, 60)
res = {k: [] for k in trainers}
for r in range(rounds):
    for name in trainers:
        res[name].append(run(name, steps))
    print(f"round {r}: " + "  ".join(f"{k} {v[-1]:.3f} ms" for k, v in res.items()), flush=True)
print(f"[ab {what}] median ms/step over {rounds} rounds of {steps} steps: " +
      "  ".join(f"{k} {statistics.median(v):.3f}" for k, v in res.items()))
