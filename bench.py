#!/usr/bin/env python
"""Headline benchmark: CLIP ViT-B/16 adapter fine-tune step (BASELINE.json configs[1]) in images/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over one batch: both frozen towers forward, adapters, projections,
global contrastive loss, adapter-only backward, grad-norm clip + AdamW (trainer.py:73-99), 256 synthetic
image/caption pairs per GPU (weak scaling).  Prints ONE JSON line (rank 0).

  value     images/s with the batch resident in HBM (CUDA-event timed, max over ranks)
  e2e       images/s through the public API (DevicePrefetcher + CLIPAdapterTrainer.training_step) from pinned
            HOST buffers, H2D of every batch and a D2H read of every loss inside the timed region
  roofline  the dominant kernel (tcgen05 dense-layer GEMM): algorithmic FLOPs of its launches / their CUDA-event
            time inside an instrumented step, against the measured bf16 peak of MEASURED_PEAKS.json
  cpu_baseline  the oracle port of the reference path (fp32 PyTorch on the host cores), bounded sample

--impl reference times that CPU path alone (the reference is pure Python and /root/reference does not travel
to the GPU box, so the arm runs the oracle port: kind "port").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("HF_HUB_OFFLINE", "1")
os.environ.setdefault("TOKENIZERS_PARALLELISM", "false")

MODEL = "openai/clip-vit-base-patch16"
BATCH = 256
ADAPTER_KIND = "clip_adapter"
METRIC = "adapter fine-tune images/sec (ViT-B/16, bf16)"
UNIT = "images/s"


WORKLOAD = ("CLIP ViT-B/16 + bottleneck adapters (A=256), frozen backbone, Track-M train step "
            "(fwd both towers, global InfoNCE, adapter bwd, clip + AdamW)")


def _peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_step_rate(steps: int, warmup: int, batch: int = 8):
    """Oracle port of the Track-M train step (model_m.py forward + trainer.py:91-99) on the host cores, fp32."""
    import torch

    from oracle import clip_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    clip = O.build_hf_clip(MODEL, seed=0)
    sd = {k: v.detach() for k, v in clip.state_dict().items()}
    d = O.CLIP_DIMS[MODEL]
    torch.manual_seed(1)

    def mk(D, A):
        lin1, lin2, ln = torch.nn.Linear(D, A), torch.nn.Linear(A, D), torch.nn.LayerNorm(D)
        return {"down_project.weight": lin1.weight, "down_project.bias": lin1.bias, "up_project.weight": lin2.weight,
                "up_project.bias": lin2.bias, "layer_norm.weight": ln.weight, "layer_norm.bias": ln.bias}

    ta, va = mk(d.text.width, 256), mk(d.vision.width, 256)
    params = list(ta.values()) + list(va.values())
    opt = torch.optim.AdamW(params, lr=5e-5, weight_decay=0.01)
    pix, ids, mask = O.synthetic_batch(batch, seed=2)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        out = O.model_m_forward(sd, d.text.heads, d.vision.heads, ids, mask, pix, ta, va)
        opt.zero_grad()
        out["loss"].backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        _ = out["loss"].item()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return {"value": batch / med, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{steps} steps of {batch} pairs (ViT-B/16, fp32, torch {torch.__version__}, "
                      f"{cores} threads), median {med:.3f} s/step; linear in batch"}, sum(times) / len(times)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 2))
    base, mean_t = cpu_reference_step_rate(steps, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": mean_t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "global_batch": BATCH * max(1, args.gpus),
                   "parallelism": f"dp{max(1, args.gpus)}",
                   "sample": "each step is a bounded sample of the workload: 8 of the 256 pairs, same model, same step "
                             "(forward both towers, InfoNCE, adapter backward, clip_grad_norm, AdamW), fp32 on the host cores; "
                             "images/s is linear in the batch on the CPU"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.power = [], set(), []
        self.max_mhz = None
        self._stop_ev = threading.Event()

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
                getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
            }
            while not self._stop_ev.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.05)
        except Exception as e:  # NVML missing: report that instead of inventing clocks
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=2)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "sm_mhz_min": min(self.samples) if self.samples else None, "samples": len(self.samples),
                "power_w_max": max(self.power) if self.power else None}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_native_arm(args):
    import torch
    import torch.distributed as dist

    from oracle import clip_oracle as O  # weight-container builder + FLOP accounting + cpu_baseline only
    from vlm_clip_b200 import _native as N
    from vlm_clip_b200 import ops
    from vlm_clip_b200.data import DevicePrefetcher
    from vlm_clip_b200.model_m import CLIPWithAdapters
    from vlm_clip_b200.trainer import CLIPAdapterTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N.load()

    clip = O.build_hf_clip(MODEL, seed=0).to(dev)
    torch.manual_seed(1)
    model = CLIPWithAdapters(clip=clip, use_shared_adapters=False, adapter_kind=ADAPTER_KIND).to(dev)
    model.train()
    trainer = CLIPAdapterTrainer(model, train_dataloader=[None], output_dir="/tmp/vlmclip_bench_ckpt")

    K, W = args.steps, max(3, args.warmup)
    nrot = 3  # rotate three different batches so no step re-reads the previous step's inputs
    g = torch.Generator().manual_seed(100 + rank)
    host = []
    for _ in range(nrot):
        pix = torch.randn(BATCH, 3, 224, 224, generator=g).pin_memory()
        ids = torch.randint(3, 49406, (BATCH, 77), generator=g)
        ids[:, 0], ids[:, -1] = 49406, 49407
        host.append({"input_ids": ids.pin_memory(), "attention_mask": torch.ones(BATCH, 77, dtype=torch.int64).pin_memory(),
                     "pixel_values": pix})
    resident = [{k: v.to(dev) for k, v in b.items()} for b in host]
    torch.cuda.synchronize()
    ready = torch.cuda.Event()
    ready.record()  # the resident batches are complete: the frozen towers need not wait for the previous step's tail
    for b in resident:
        b["inputs_ready"] = ready

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- value: inputs resident in HBM ----------------
    for i in range(W):
        trainer.training_step(resident[i % nrot])
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    n0 = N.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        loss = trainer.training_step(resident[i % nrot])
    e1.record()
    barrier()
    n1 = N.launch_count()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop()
    ms_step = ms_total / K
    value = world * BATCH * K / (ms_total / 1e3)
    final_loss = float(loss.item())

    # ---------------- e2e: public API from pinned host buffers ----------------
    class _HostLoader:
        def __init__(self, n):
            self.n = n

        def __len__(self):
            return self.n

        def __iter__(self):
            for i in range(self.n):
                yield host[i % nrot]

    loss_host = torch.empty(K + W, dtype=torch.float32).pin_memory()
    barrier()
    it = 0
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for batch in DevicePrefetcher(_HostLoader(W + K), dev):
        if it == W:
            barrier()
            e2.record()
        l = trainer.training_step(batch)
        loss_host[it:it + 1].copy_(l.reshape(1), non_blocking=True)  # D2H read of every step's loss
        it += 1
    e3.record()
    barrier()
    ms_e2e = max_over_ranks(e2.elapsed_time(e3))
    e2e_value = world * BATCH * K / (ms_e2e / 1e3)
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())

    # ---------------- opt-in shortcut variant (reported beside the headline, never as it) ----------------
    # Track M pools the causal text tower at token 0 (the reference's BOS quirk), so the text tower on
    # that single token gives identical features, and it keeps only the CLS row of the vision tower's output
    # (model_m.py:122), so the last vision layer only needs that row after its QKV GEMM; `value` above is the DENSE
    # computation, this is the same step with both flags on and the skipped FLOPs taken out of its TFLOP count
    # (SURVEY.md 8d).
    model.text_token0_only = True
    model.vision_cls_only_last_layer = True
    Ks = max(3, K // 4)
    for i in range(3):
        trainer.training_step(resident[i % nrot])
    barrier()
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record()
    for i in range(Ks):
        loss_s = trainer.training_step(resident[i % nrot])
    e5.record()
    barrier()
    ms_short = max_over_ranks(e4.elapsed_time(e5)) / Ks
    model.text_token0_only = False
    model.vision_cls_only_last_layer = False

    # ---------------- roofline of the dominant kernel (instrumented step) ----------------
    # (towers serialised on one stream for this step only, so that a launch's event pair brackets that kernel alone)
    overlap = model.overlap_towers
    model.overlap_towers = False
    ops.PROFILE = {"gemm": []}
    trainer.training_step(resident[0])
    torch.cuda.synchronize()
    model.overlap_towers = overlap
    gemm_ms = sum(a.elapsed_time(b) for a, b, _ in ops.PROFILE["gemm"])
    gemm_fl = sum(f for _, _, f in ops.PROFILE["gemm"])
    n_gemm = len(ops.PROFILE["gemm"])
    ops.PROFILE = None
    peaks, peak_kind = _peaks()
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    achieved = gemm_fl / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    fl = O.flops_per_pair(MODEL)
    step_tf = fl["pair"] * BATCH / 1e12
    traffic = None
    tf = ROOT / "profiles" / "gemm_traffic.json"
    if tf.exists():
        traffic = json.loads(tf.read_text()).get("dram_bytes_per_launch")

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {
            "workload": WORKLOAD,
            "batch_per_gpu": BATCH, "global_batch": BATCH * world, "parallelism": f"dp{world}",
            "init": "random (seed 0), no checkpoints offline",
            "l2": f"3 rotating input batches; {BATCH * 3 * 224 * 224 * 4 / 1e6:.0f} MB pixel batch and >1 GB of activations per step exceed the 126 MB L2",
            "algorithmic_tflop_per_step_per_gpu": step_tf,
            "step_tflops_per_gpu": step_tf / (ms_step / 1e3),
            "step_frac_of_bf16_sustained_peak": step_tf / (ms_step / 1e3) / peak_tf,
            "final_loss": final_loss,
            "streams": ("frozen towers on two private CUDA streams (they start on the input-ready event, so they overlap the "
                        "previous step's adapter backward / AdamW); adapters, loss, backward, optimizer on the main stream")
            if model.overlap_towers else "single stream",
        },
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / K},
        "gpu_launches": int(n1 - n0),
        "shortcut_variant": {
            "what": "text tower evaluated on token 0 only (result-identical for Track M's BOS pooling under the causal mask) and "
                    "last vision layer evaluated for the CLS row only after its QKV GEMM; opt-in flags "
                    "model.text_token0_only / model.vision_cls_only_last_layer, OFF for every other number in this line",
            "value": world * BATCH / (ms_short / 1e3), "unit": UNIT, "ms_per_step": ms_short, "steps": Ks,
            "executed_tflop_per_step_per_gpu": ((fl["image"] - _cls_only_skipped_flops(MODEL)) * BATCH
                                                + fl["caption"] * BATCH / 77.0) / 1e12,
        },
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved / peak_tf if peak_tf else None, "traffic": traffic,
                     "kernel": "gemm_bf16_tn_kernel (tcgen05)", "launches_per_step": n_gemm,
                     "avg_launch_ms": gemm_ms / max(1, n_gemm), "share_of_step": gemm_ms / ms_step,
                     "peak_source": f"{peak_kind} bf16_tflops_sustained (kernel timed inside a step)"},
    }
    if world == 1 and not args.no_full_finetune:
        line["full_finetune_variant"] = _full_finetune_leg(dev, fl, peak_tf)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base, _ = cpu_reference_step_rate(steps=3, warmup=1)
        line["cpu_baseline"] = base
    if rank == 0:
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


def _full_finetune_leg(dev, fl, peak_tf, steps: int = 5):
    """BASELINE config 5 beside the headline: the same Track-M step with the adapters disabled and EVERY CLIP parameter
    trainable (`freeze_clip=False`): towers forward from the live fp32 weights, hand-written backward through both
    towers (dgrad / wgrad on the tcgen05 GEMM, tensor-core attention backward), clip + AdamW over 151 M parameters.
    Its own model instance; a failure here is reported in the record and never touches the headline numbers."""
    import torch

    from oracle import clip_oracle as O
    from vlm_clip_b200 import _native as N
    from vlm_clip_b200.model_m import CLIPWithAdapters
    from vlm_clip_b200.trainer import CLIPAdapterTrainer

    try:
        clip = O.build_hf_clip(MODEL, seed=0).to(dev)
        model = CLIPWithAdapters(clip=clip, freeze_clip=False, use_text_adapter=False, use_vision_adapter=False,
                                 use_shared_adapters=False).to(dev)
        model.train()
        g = torch.Generator().manual_seed(7)
        ids = torch.randint(3, 49406, (BATCH, 77), generator=g)
        ids[:, 0], ids[:, -1] = 49406, 49407
        batch = {"input_ids": ids.to(dev), "attention_mask": torch.ones(BATCH, 77, dtype=torch.int64, device=dev),
                 "pixel_values": torch.randn(BATCH, 3, 224, 224, generator=g).to(dev)}
        tr = CLIPAdapterTrainer(model, [batch], learning_rate=1e-7, output_dir="/tmp/vlmclip_bench_ft", trainable="all")
        for _ in range(3):
            tr.training_step(batch)
        torch.cuda.synchronize()
        n0 = N.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = tr.training_step(batch)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        tflop = 3.0 * fl["pair"] * BATCH / 1e12  # forward + input gradients + weight gradients
        return {"what": "BASELINE config 5: full fine-tune, adapters disabled, all CLIP parameters trainable "
                        "(CLIPWithAdapters(freeze_clip=False), trainer(trainable='all')); same batch of 256 pairs",
                "value": BATCH / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
                "trainable_parameters": sum(p.numel() for p in tr.trainable_params),
                "algorithmic_tflop_per_step": tflop, "step_tflops": tflop / (ms / 1e3),
                "step_frac_of_bf16_sustained_peak": tflop / (ms / 1e3) / peak_tf,
                "gpu_launches_per_step": int((N.launch_count() - n0) / steps), "final_loss": float(loss.item()),
                "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
    except Exception as e:  # never let the comparison config break the headline line
        return {"error": f"{type(e).__name__}: {e}"}


def _cls_only_skipped_flops(model_name: str) -> float:
    """FLOPs of the last vision layer that the CLS-only evaluation does not execute (per image)."""
    from oracle import clip_oracle as O

    v = O.CLIP_DIMS[model_name].vision
    S, D, F = v.seq, v.width, v.mlp
    return float((S - 1) * (2 * D * D + 4 * D * F) + 4 * (S - 1) * S * D)


_JSON_OUT = None


def _emit(line: dict) -> None:
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # stdout carries exactly ONE JSON line: everything else a library prints there (NCCL's version banner under
    # torchrun, for one) is sent to stderr by pointing fd 1 at fd 2 and keeping a private handle on the real stdout.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)  # ~1.2 s timed: long enough to sit at the sustained (power-capped) clock
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full-finetune", action="store_true", help="skip the config-5 (full fine-tune) comparison leg")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3"],
                    help="cfg2 (default, the headline): ViT-B/16 + bottleneck adapters, 256 pairs per GPU.  cfg3 (BASELINE "
                         "configs[2], not a headline line): ViT-L/14 + PE-CLIP adapters, 512 pairs per GPU = global batch 4096 "
                         "on 8 GPUs, global contrastive loss over the all-gathered embeddings")
    args = ap.parse_args()
    if args.workload == "cfg3":
        global MODEL, BATCH, METRIC, WORKLOAD, ADAPTER_KIND
        MODEL, BATCH, ADAPTER_KIND = "openai/clip-vit-large-patch14", 512, "peclip"
        METRIC = "adapter fine-tune images/sec (ViT-L/14 + PE-CLIP adapters, bf16)"
        WORKLOAD = ("CLIP ViT-L/14 + PE-CLIP adapters (A=256), frozen backbone, Track-M train step with the global contrastive "
                    "loss over all-gathered embeddings (BASELINE config 3)")
        args.no_full_finetune = True
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_native_arm(args)


if __name__ == "__main__":
    main()
