"""Runs the UNMODIFIED reference (staged by oracle/stage_reference.py) on the host cores — TEST INFRASTRUCTURE ONLY.

Used by bench.py's `--impl reference` arm and `cpu_baseline` leg.  Shim (a monkey-patch in this process, SURVEY.md §8c
shim 1; no reference file is touched): `CLIPModel.from_pretrained` returns a seeded random-init `CLIPModel` of the named
checkpoint's dimensions (there are no weights offline) and `CLIPProcessor.from_pretrained` a stub.  Everything else —
`model_m.CLIPWithAdapters.forward`, `trainer.CLIPAdapterTrainer.train`'s loop body (trainer.py:73-103: forward, backward,
clip_grad_norm_, AdamW.step, scheduler.step, loss.item) — is the reference's own code, timed from the outside by a
data loader that records when each batch is requested.
"""
from __future__ import annotations

import os
import statistics
import sys
import time
import types
from pathlib import Path

STAGED = Path(__file__).resolve().parent / "_ref" / "reference"


def available() -> bool:
    return (STAGED / "model_m.py").exists() and (STAGED / "trainer.py").exists()


class _TimedLoader:
    """A list-like loader; t[i] is the moment batch i was requested, t[-1] the moment the epoch ended."""

    def __init__(self, batches):
        self.batches = batches
        self.t = []

    def __len__(self):
        return len(self.batches)

    def __iter__(self):
        for b in self.batches:
            self.t.append(time.perf_counter())
            yield b
        self.t.append(time.perf_counter())


def reference_step_rate(model_name: str, steps: int, warmup: int, batch: int, seed_model: int = 0, threads=None):
    """images/s of the reference's own trainer loop over `warmup + steps` synthetic batches of `batch` pairs (fp32, CPU)."""
    import torch
    from transformers import CLIPModel, CLIPProcessor

    from . import clip_oracle as O

    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    CLIPModel.from_pretrained = staticmethod(lambda name, *a, **k: O.build_hf_clip(name, seed=seed_model))
    CLIPProcessor.from_pretrained = staticmethod(lambda name, *a, **k: types.SimpleNamespace(name=name))
    if str(STAGED) not in sys.path:
        sys.path.insert(0, str(STAGED))
    import model_m as ref_model_m  # the reference's module, unmodified
    import trainer as ref_trainer

    torch.manual_seed(1)
    model = ref_model_m.CLIPWithAdapters(clip_model_name=model_name, use_shared_adapters=False)
    batches = []
    for i in range(warmup + steps):
        pix, ids, mask = O.synthetic_batch(batch, seed=2 + i)
        batches.append({"input_ids": ids, "attention_mask": mask, "pixel_values": pix})
    loader = _TimedLoader(batches)
    tr = ref_trainer.CLIPAdapterTrainer(model, loader, output_dir="/tmp/vlmclip_ref_arm")
    tr.train(num_epochs=1, save_every=10 ** 9)
    dt = [b - a for a, b in zip(loader.t[:-1], loader.t[1:])][warmup:]
    med = statistics.median(dt)
    return {"value": batch / med, "unit": "images/s", "cores": cores, "kind": "reference",
            "sample": f"{steps} steps of {batch} pairs through the reference's own model_m.CLIPWithAdapters + "
                      f"trainer.CLIPAdapterTrainer.train loop ({model_name}, random init, fp32, torch {torch.__version__}, "
                      f"{cores} threads), median {med:.3f} s/step; images/s is linear in the batch on the CPU"}, sum(dt) / len(dt)
