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
            time inside instrumented steps (median of 10), against the measured bf16 peaks of MEASURED_PEAKS.json
  cpu_baseline  the reference path on the host cores (fp32), bounded sample: the unmodified reference modules staged
            into oracle/_ref/ (kind "reference"), or the oracle port when they did not travel (kind "port")

--impl reference times that CPU path alone.
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
    """The reference path on the host cores, fp32: the UNMODIFIED reference modules (model_m.CLIPWithAdapters +
    trainer.CLIPAdapterTrainer.train, staged into oracle/_ref/ by oracle/stage_reference.py; kind "reference") when they
    travelled with the snapshot, else the oracle port of the same step (kind "port")."""
    from oracle import ref_harness as H

    if H.available():
        return H.reference_step_rate(MODEL, steps, warmup, batch)
    return cpu_port_step_rate(steps, warmup, batch)


def cpu_port_step_rate(steps: int, warmup: int, batch: int = 8):
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

    from vlm_clip_b200 import _native as N
    from vlm_clip_b200 import ops
    from vlm_clip_b200.configs import flops_per_pair, random_init_clip
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

    clip = random_init_clip(MODEL, seed=0).to(dev)
    torch.manual_seed(1)
    model = CLIPWithAdapters(clip=clip, use_shared_adapters=False, adapter_kind=ADAPTER_KIND).to(dev)
    model.train()
    use_graph = not args.no_graph
    trainer = CLIPAdapterTrainer(model, train_dataloader=[None], output_dir="/tmp/vlmclip_bench_ckpt", cuda_graph=use_graph)
    graph_note = None

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

    def launches_now() -> int:
        # kernels of the library: counted at enqueue time, plus the ones every graph replay re-launches
        return int(N.launch_count()) + trainer.graph_replays * trainer.graph_launches_per_step

    class _HostLoader:
        def __init__(self, batches, n):
            self.batches, self.n = batches, n

        def __len__(self):
            return self.n

        def __iter__(self):
            for i in range(self.n):
                yield self.batches[i % len(self.batches)]

    def timed_resident(k_steps: int, w_steps: int):
        """(ms total max over ranks, host ms spent enqueueing the k steps, launches, last loss)"""
        for i in range(w_steps):
            trainer.training_step(resident[i % nrot])
        barrier()
        n0 = launches_now()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h0 = time.perf_counter()
        loss = None
        for i in range(k_steps):
            loss = trainer.training_step(resident[i % nrot])
        h1 = time.perf_counter()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), (h1 - h0) * 1e3, launches_now() - n0, loss

    def timed_e2e(batches, k_steps: int, w_steps: int):
        """Public API from pinned HOST buffers: DevicePrefetcher (H2D of every batch) + training_step + D2H of every loss."""
        loss_host = torch.empty(k_steps + w_steps, dtype=torch.float32).pin_memory()
        barrier()
        it = 0
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0 = time.perf_counter()
        for batch in DevicePrefetcher(_HostLoader(batches, w_steps + k_steps), dev):
            if it == w_steps:
                barrier()
                e2.record()
                h0 = time.perf_counter()
            l = trainer.training_step(batch)
            loss_host[it:it + 1].copy_(l.reshape(1), non_blocking=True)  # D2H read of every step's loss
            it += 1
        h1 = time.perf_counter()
        e3.record()
        barrier()
        return max_over_ranks(e2.elapsed_time(e3)), (h1 - h0) * 1e3

    # ---------------- settle: bring the board to its sustained, power-capped state before anything is timed -----------
    # (the driver's default K = 20 is a 0.25 s region; without this it runs at burst clocks while the roofline fraction is
    # quoted against the SUSTAINED peak)
    t_settle = time.perf_counter()
    n_settle = 0
    try:
        while True:
            for _ in range(8):  # every rank runs the SAME number of steps: each one contains collectives
                trainer.training_step(resident[n_settle % nrot])
                n_settle += 1
            torch.cuda.synchronize()
            # the slowest rank's clock decides for everybody
            if max_over_ranks((time.perf_counter() - t_settle) * 1e3) >= args.settle_s * 1e3 and \
                    n_settle >= trainer.graph_warmup_steps + 2:
                break
    except Exception as e:  # a capture failure must not cost the round its benchmark line: same kernels, eager launches
        if not use_graph:
            raise
        graph_note = f"capture failed, eager launches used: {type(e).__name__}: {e}"
        sys.stderr.write(graph_note + "\n")
        torch.cuda.synchronize()
        trainer.cuda_graph = False
        use_graph = False

    # ---------------- value: inputs resident in HBM ----------------
    sampler = ClockSampler(local)
    sampler.start()
    ms_total, host_ms, n_launch, loss = timed_resident(K, W)
    clocks = sampler.stop()
    # the trainable half of the step (G_H: final LN, adapters, projections, all-gather, loss, LSE exchange, backward,
    # gradient all-reduce, clip + AdamW) timed by itself over a few extra steps: it runs UNDER the next step's towers,
    # so this is its own duration (incl. waiting for SMs and for the slowest rank's collectives), not exposed time
    tail = None
    if use_graph:
        trainer.tail_events = []
        for i in range(10):
            trainer.training_step(resident[i % nrot])
        torch.cuda.synchronize()
        tms = sorted(a.elapsed_time(b) for a, b in trainer.tail_events)
        trainer.tail_events = None
        if tms:
            tail = {"what": "duration of the trainable half of a step (graph G_H) on the caller's stream, overlapped with the "
                            "next step's towers: heads, feature all-gather, loss strips, LSE all-gather, backward, gradient "
                            "all-reduce, clip + AdamW", "ms_median": statistics.median(tms), "ms_max": tms[-1],
                    "ms_min": tms[0], "steps": len(tms)}
    ms_step = ms_total / K
    value = world * BATCH * K / (ms_total / 1e3)
    final_loss = float(loss.item())

    # ---------------- e2e: public API from pinned host buffers ----------------
    ms_e2e, host_ms_e2e = timed_e2e(host, K, W)
    e2e_value = world * BATCH * K / (ms_e2e / 1e3)
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())
    # one isolated pinned-host -> device copy of a pixel batch: what this box's PCIe path delivers
    pb = host[0]["pixel_values"]
    dst = torch.empty_like(resident[0]["pixel_values"])
    dst.copy_(pb, non_blocking=True)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(3):
        dst.copy_(pb, non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    h2d_gbs = 3 * pb.numel() * pb.element_size() / (c0.elapsed_time(c1) / 1e3) / 1e9
    del dst

    # ---------------- e2e from uint8 frames (a quarter of the H2D bytes; fused preprocessing on the GPU) -------------
    u8_variant = None
    if not args.lean:
        host_u8 = []
        for b in host:
            fr = torch.randint(0, 256, (BATCH, 224, 224, 3), generator=g, dtype=torch.uint8).pin_memory()
            host_u8.append({"input_ids": b["input_ids"], "attention_mask": b["attention_mask"], "pixel_values": fr})
        Ku = max(5, K // 2)
        ms_u8, _ = timed_e2e(host_u8, Ku, trainer.graph_warmup_steps + 3)
        u8_variant = {
            "what": "same public API fed decoded uint8 frames [B, 224, 224, 3] (resize / scale / normalise fused into "
                    "the patch extraction on the GPU) instead of fp32 pixel_values: a quarter of the H2D bytes",
            "value": world * BATCH / (ms_u8 / Ku / 1e3), "unit": UNIT, "ms_per_step": ms_u8 / Ku,
            "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in host_u8[0].values()), "steps": Ku}
        del host_u8

    # ---------------- the same step with eager launches (no CUDA graph), for comparison ----------------
    eager = None
    if use_graph and not args.lean:
        trainer.cuda_graph = False
        Ke = K
        ms_eager, host_eager, _, _ = timed_resident(Ke, W)
        ms_eager_e2e, host_eager_e2e = timed_e2e(host, Ke, W)
        trainer.cuda_graph = True
        eager = {"what": "same step, kernels enqueued one by one (two native tower calls + ~60 interpreter-level ops per step)",
                 "ms_per_step": ms_eager / Ke, "host_enqueue_ms_per_step": host_eager / Ke,
                 "e2e_ms_per_step": ms_eager_e2e / Ke, "e2e_host_enqueue_ms_per_step": host_eager_e2e / Ke, "steps": Ke}

    # ---------------- opt-in shortcut variant (reported beside the headline, never as it) ----------------
    # Track M pools the causal text tower at token 0 (the reference's BOS quirk), so the text tower on
    # that single token gives identical features, and it keeps only the CLS row of the vision tower's output
    # (model_m.py:122), so the last vision layer only needs that row after its QKV GEMM; `value` above is the DENSE
    # computation, this is the same step with both flags on and the skipped FLOPs taken out of its TFLOP count
    # (SURVEY.md 8d).
    shortcut = None
    if not args.lean:
        model.text_token0_only = True
        model.vision_cls_only_last_layer = True
        Ks = max(5, K // 4)
        ms_short_total, _, _, _ = timed_resident(Ks, trainer.graph_warmup_steps + 3)
        ms_short = ms_short_total / Ks
        model.text_token0_only = False
        model.vision_cls_only_last_layer = False
        fl_ = flops_per_pair(MODEL)
        shortcut = {
            "what": "text tower evaluated on token 0 only (result-identical for Track M's BOS pooling under the causal mask) and "
                    "last vision layer evaluated for the CLS row only after its QKV GEMM; opt-in flags "
                    "model.text_token0_only / model.vision_cls_only_last_layer, OFF for every other number in this line",
            "value": world * BATCH / (ms_short / 1e3), "unit": UNIT, "ms_per_step": ms_short, "steps": Ks,
            "executed_tflop_per_step_per_gpu": ((fl_["image"] - _cls_only_skipped_flops(MODEL)) * BATCH
                                                + fl_["caption"] * BATCH / 77.0) / 1e12,
        }

    # ---------------- roofline of the dominant kernel (instrumented, eager steps) ----------------
    # (towers serialised on one stream, so that a launch's event pair brackets that kernel alone; every instrumented
    # step is a whole train step, its GEMM launches are timed one by one)
    overlap = model.overlap_towers
    model.overlap_towers = False
    trainer.cuda_graph = False
    sampler2 = ClockSampler(local)
    sampler2.start()
    per_step = []
    n_gemm = 0
    for r in range(max(1, 2 if args.lean else args.roofline_steps)):
        ops.PROFILE = {"gemm": []}
        trainer.training_step(resident[r % nrot])
        torch.cuda.synchronize()
        rec = ops.PROFILE["gemm"]
        ops.PROFILE = None
        g_ms = sum(a.elapsed_time(b) for a, b, _ in rec)
        g_fl = sum(f for _, _, f in rec)
        n_gemm = len(rec)
        per_step.append((g_fl / (g_ms / 1e3) / 1e12, g_ms))
    clocks_roof = sampler2.stop()
    model.overlap_towers = overlap
    trainer.cuda_graph = use_graph
    peaks, peak_kind = _peaks()
    peak_sus = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    peak_burst = float(peaks["bf16_tflops"])
    ach = sorted(a for a, _ in per_step)
    achieved = statistics.median(ach)
    gemm_ms = statistics.median(m for _, m in per_step)
    fl = flops_per_pair(MODEL)
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
            "residual_stream": model._backbone().residual,
            "launch": ("whole step captured as TWO CUDA graphs per batch signature and replayed: G_T (both frozen towers on two "
                       "branches -> fp32 pooled rows) on its own stream, G_H (final LN, adapters, projections, loss, backward, "
                       "NCCL, clip + AdamW) on the caller's; G_T of step k+1 runs under G_H of step k (the towers do not depend "
                       "on the optimizer); per step the host copies the batch into G_T's input slot (device to device) and "
                       "launches two graphs") if use_graph else
                      ("eager launches" + (f" ({graph_note})" if graph_note else "")),
            "settle": f"{n_settle} untimed steps ({time.perf_counter() - t_settle:.1f} s incl. graph capture) before the {W} warm-up "
                      "steps, so that the timed region runs at the sustained power-capped clock",
            "algorithmic_tflop_per_step_per_gpu": step_tf,
            "step_tflops_per_gpu": step_tf / (ms_step / 1e3),
            "step_frac_of_bf16_sustained_peak": step_tf / (ms_step / 1e3) / peak_sus,
            "step_frac_of_bf16_burst_peak": step_tf / (ms_step / 1e3) / peak_burst,
            "host_enqueue_ms_per_step": host_ms / K,
            "final_loss": final_loss,
            "streams": ("frozen towers on two private CUDA streams; adapters, loss, backward, optimizer on the main stream")
            if model.overlap_towers else "single stream",
        },
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / K, "host_enqueue_ms_per_step": host_ms_e2e / K,
                "h2d_gbs_needed_to_hide_copy": h2d / (ms_e2e / K / 1e3) / 1e9, "h2d_gbs_isolated_copy": h2d_gbs,
                "uint8_frames_variant": u8_variant},
        "gpu_launches": int(n_launch),
        "tail": tail,
        "shortcut_variant": shortcut,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_sus, "unit": "TFLOP/s",
                     "frac": achieved / peak_sus if peak_sus else None, "traffic": traffic,
                     "kernel": "gemm_bf16_tn_kernel (tcgen05)", "launches_per_step": n_gemm,
                     "avg_launch_ms": gemm_ms / max(1, n_gemm), "share_of_step": gemm_ms / ms_step,
                     "instrumented_steps": len(per_step), "achieved_min": ach[0], "achieved_max": ach[-1],
                     "frac_of_burst_peak": achieved / peak_burst if peak_burst else None, "peak_burst": peak_burst,
                     "sm_mhz_median_during_instrumented_steps": clocks_roof.get("sm_mhz"),
                     "peak_source": f"{peak_kind} bf16_tflops_sustained (kernel timed inside a step); burst = bf16_tflops"},
    }
    if eager is not None:
        line["eager_variant"] = eager
    if world == 1 and not args.no_full_finetune:
        line["full_finetune_variant"] = _full_finetune_leg(dev, fl, peak_sus)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base, _ = cpu_reference_step_rate(steps=3, warmup=1)
        line["cpu_baseline"] = base
    if rank == 0:
        _emit(line)
    _shutdown(world, trainer)


def _shutdown(world: int, trainer=None):
    """Orderly end of a multi-rank run.  The step's collectives live inside CUDA graphs: the graphs go first (NCCL cannot
    tear a communicator down under them - `destroy_process_group()` hung for the full 900 s limit of the first 2-GPU
    run), then a barrier, then the process leaves without running NCCL's destructors."""
    import torch
    import torch.distributed as dist

    if trainer is not None:
        trainer.release_graphs()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stderr.flush()
        os._exit(0)


def run_finetune_arm(args):
    """BASELINE config 5 as its own bench line (not the headline): full fine-tune of ViT-B/16 (adapters disabled, all
    151 M CLIP parameters trainable), 256 pairs per GPU, data parallel with the gradient arena all-reduced in per-layer
    buckets that start inside the backward pass (dist.BucketedGradAllReduce)."""
    import torch
    import torch.distributed as dist

    from vlm_clip_b200 import _native as N
    from vlm_clip_b200.configs import flops_per_pair, random_init_clip
    from vlm_clip_b200.model_m import CLIPWithAdapters
    from vlm_clip_b200.trainer import CLIPAdapterTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N.load()
    clip = random_init_clip(MODEL, seed=0).to(dev)
    model = CLIPWithAdapters(clip=clip, freeze_clip=False, use_text_adapter=False, use_vision_adapter=False,
                             use_shared_adapters=False).to(dev)
    model.train()
    g = torch.Generator().manual_seed(7 + rank)
    batches = []
    for _ in range(2):
        ids = torch.randint(3, 49406, (BATCH, 77), generator=g)
        ids[:, 0], ids[:, -1] = (torch.arange(BATCH) * 37 + rank) % 49000, 49407
        batches.append({"input_ids": ids.to(dev), "attention_mask": torch.ones(BATCH, 77, dtype=torch.int64, device=dev),
                        "pixel_values": torch.randn(BATCH, 3, 224, 224, generator=g).to(dev)})
    tr = CLIPAdapterTrainer(model, batches, learning_rate=1e-7, output_dir="/tmp/vlmclip_bench_ft", trainable="all")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(k_steps, w_steps):
        for i in range(w_steps):
            tr.training_step(batches[i % 2])
        barrier()
        n0 = N.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss = None
        for i in range(k_steps):
            loss = tr.training_step(batches[i % 2])
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, int(N.launch_count() - n0), float(loss.item())

    K, W = args.steps, max(3, args.warmup)
    sampler = ClockSampler(local)
    sampler.start()
    ms_total, n_launch, final_loss = timed(K, W)
    clocks = sampler.stop()
    ms_step = ms_total / K
    # the same step without any gradient collective (replicas drift apart: timing only), to size the collective's share
    share = None
    if world > 1:
        import vlm_clip_b200.trainer as T

        saved_ar, saved_b = T.allreduce_sum_, tr._grad_buckets
        T.allreduce_sum_ = lambda t, group=None: t
        tr._grad_buckets = lambda: None
        model._finetune_towers().layer_grad_sink = None
        ms_nocoll, _, _ = timed(max(3, K // 2), 2)
        T.allreduce_sum_, tr._grad_buckets = saved_ar, saved_b
        ms_nocoll /= max(3, K // 2)
        share = {"ms_per_step_without_gradient_allreduce": ms_nocoll, "exposed_collective_ms": ms_step - ms_nocoll,
                 "exposed_share_of_step": (ms_step - ms_nocoll) / ms_step}
    fl = flops_per_pair(MODEL)
    peaks, _ = _peaks()
    peak_sus = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    tflop = 3.0 * fl["pair"] * BATCH / 1e12
    n_par = sum(p.numel() for p in tr.trainable_params)
    line = {
        "metric": "full fine-tune images/sec (ViT-B/16, bf16)", "value": world * BATCH / (ms_step / 1e3), "unit": UNIT,
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "BASELINE config 5: CLIP ViT-B/16 full fine-tune (adapters disabled, every CLIP parameter "
                               "trainable), Track-M step with hand-written tower backward, clip + AdamW over one arena",
                   "batch_per_gpu": BATCH, "global_batch": BATCH * world, "parallelism": f"dp{world}",
                   "trainable_parameters": n_par, "gradient_bytes_per_step": 4 * n_par,
                   "allreduce": f"{tr.last_allreduce_buckets} NCCL all-reduce calls per step: one per encoder layer issued from "
                                "inside the backward (last layers first) + the remaining arena ranges" if world > 1 else "none",
                   "algorithmic_tflop_per_step_per_gpu": tflop, "step_tflops_per_gpu": tflop / (ms_step / 1e3),
                   "step_frac_of_bf16_sustained_peak": tflop / (ms_step / 1e3) / peak_sus, "final_loss": final_loss,
                   "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30},
        "clocks": clocks, "gpu_launches": n_launch,
    }
    if share is not None:
        line["collective"] = share
    if rank == 0:
        _emit(line)
    _shutdown(world)


def _full_finetune_leg(dev, fl, peak_tf, steps: int = 5):
    """BASELINE config 5 beside the headline: the same Track-M step with the adapters disabled and EVERY CLIP parameter
    trainable (`freeze_clip=False`): towers forward from the live fp32 weights, hand-written backward through both
    towers (dgrad / wgrad on the tcgen05 GEMM, tensor-core attention backward), clip + AdamW over 151 M parameters.
    Its own model instance; a failure here is reported in the record and never touches the headline numbers."""
    import torch

    from vlm_clip_b200 import _native as N
    from vlm_clip_b200.configs import random_init_clip
    from vlm_clip_b200.model_m import CLIPWithAdapters
    from vlm_clip_b200.trainer import CLIPAdapterTrainer

    try:
        clip = random_init_clip(MODEL, seed=0).to(dev)
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
    from vlm_clip_b200.configs import CLIP_DIMS

    D, _, _, F, S = CLIP_DIMS[model_name][0]
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
    ap.add_argument("--lean", action="store_true", help="skip the side legs (eager / uint8 / shortcut variants, 2 roofline steps)")
    ap.add_argument("--no-graph", action="store_true", help="enqueue the step's kernels one by one instead of replaying a CUDA graph")
    ap.add_argument("--settle-s", type=float, default=1.5, help="seconds of untimed steps before the warm-up (sustained clocks)")
    ap.add_argument("--roofline-steps", type=int, default=10, help="instrumented steps behind the roofline object")
    ap.add_argument("--no-full-finetune", action="store_true", help="skip the config-5 (full fine-tune) comparison leg")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg5"],
                    help="cfg2 (default, the headline): ViT-B/16 + bottleneck adapters, 256 pairs per GPU.  cfg3 (BASELINE "
                         "configs[2], not a headline line): ViT-L/14 + PE-CLIP adapters, 512 pairs per GPU = global batch 4096 "
                         "on 8 GPUs, global contrastive loss over the all-gathered embeddings.  cfg5 (BASELINE configs[4], not a headline line): "
                         "full fine-tune of ViT-B/16, 256 pairs per GPU, bucketed gradient all-reduce overlapped with the backward")
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
    elif args.workload == "cfg5":
        run_finetune_arm(args)
    else:
        run_native_arm(args)


if __name__ == "__main__":
    main()
