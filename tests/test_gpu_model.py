"""End-to-end parity on the GPU: the native towers / Track-M model / trainer step against the fp32 oracle on the
same seeded weights and inputs.

Tolerances.  BASELINE.json's north_star asks for bf16 logits within 1e-2 relative, fp32 loss within 1e-4 and
identical argmax; tests/test_gpu_parity_fullsize.py holds exactly those numbers at the BASELINE configs (ViT-B/16 with
256 pairs: logits 7.0e-3, loss 1.2e-5; ViT-L/14 at full depth) next to torch's own bf16 autocast as a yardstick.  The
tests in THIS file run ViT-B/32 with 4-8 pairs, where the same feature error (2e-3 image / 6e-3 text, the floor set by
the bf16 operand roundings, oracle/emulate_bf16.py; the residual stream is two-term and no longer adds to it) is averaged
over fewer rows and spread over fewer, smaller logits (random-init features are nearly orthogonal, |cos| ~ 0.03).  The
bounds below are those small-batch values with ~1.5x margin; the loss kernel itself meets 1e-4 on equal features
(test_loss_kernel_on_oracle_features, tests/test_gpu_kernels.py::test_clip_loss)."""
FEAT_TOL = 1e-2     # pooled features / hidden states, relative L2
LOGIT_TOL = 2e-2    # logits of 4-8 near-orthogonal random-init pairs, relative L2
LOSS_TOL = 5e-4     # end-to-end loss through the bf16 towers at the random-init logit scale, absolute
import math

import pytest
import torch

from oracle import clip_oracle as O

pytestmark = pytest.mark.gpu
bf16, f32 = torch.bfloat16, torch.float32
B32 = "openai/clip-vit-base-patch32"


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def clip_b32(cuda):
    m = O.build_hf_clip(B32, seed=0).to(cuda)
    for p in m.parameters():
        p.requires_grad_(False)
    return m


@pytest.fixture(scope="module")
def sd_b32(clip_b32):
    return {k: v.detach() for k, v in clip_b32.state_dict().items()}


@pytest.mark.parametrize("fold", [True, False])
def test_towers_match_oracle(cuda, clip_b32, sd_b32, fold):
    from vlm_clip_b200.towers import NativeClipTowers

    tw = NativeClipTowers(clip_b32, cuda, fold_ln=fold)
    pix, ids, mask = O.synthetic_batch(5, seed=3)
    mask[1, 30:] = 0
    mask[3, 5:] = 0
    pix, ids, mask = pix.to(cuda), ids.to(cuda), mask.to(cuda)
    v = tw.vision_hidden(pix).view(5, 50, 768)
    t = tw.text_hidden(ids, mask).view(5, 77, 512)
    with torch.no_grad():
        vo = O.vision_tower(sd_b32, pix, 12)
        to = O.text_tower(sd_b32, ids, mask, 8)
    # fold=False is the explicit-LayerNorm path that exists to bound the fold's error; it keeps the one-plane bf16
    # residual stream (9e-3 on the hidden states, round 1's figure), the product path (fold=True) the two-term one
    tol = FEAT_TOL if fold else 2e-2
    assert _rel(v, vo) < tol, _rel(v, vo)
    assert _rel(t, to) < tol, _rel(t, to)
    fi = tw.image_features(pix)
    ft = tw.text_features(ids, mask)
    with torch.no_grad():
        assert _rel(fi, O.hf_pooled_image_features(sd_b32, pix, 12)) < tol
        assert _rel(ft, O.hf_pooled_text_features(sd_b32, ids, mask, 8)) < tol


def _adapters_sd(model):
    return ({k: v.detach().clone() for k, v in model.text_adapter.state_dict().items()},
            {k: v.detach().clone() for k, v in model.vision_adapter.state_dict().items()})


def _make_model(cuda, clip, seed=1):
    from vlm_clip_b200.model_m import CLIPWithAdapters

    torch.manual_seed(seed)
    m = CLIPWithAdapters(clip=clip, use_shared_adapters=False).to(cuda)
    return m


@pytest.mark.parametrize("vary_tok0,scale", [(False, None), (True, None), (True, 100.0)])
def test_model_m_forward_backward(cuda, clip_b32, sd_b32, vary_tok0, scale):
    """scale=None keeps the random-init logit_scale (exp = 14.3); scale=100 is the pretrained value, where the
    softmax is peaked and the adapter gradient is well conditioned.  In the near-uniform random-init regime the
    batch-summed adapter gradient is a small residual of cancelling per-sample terms (exactly so when every caption
    shares token 0: sum_i P_ij ~ 1), so there it is only checked for direction."""
    sd = dict(sd_b32)
    old = clip_b32.logit_scale.data.clone()
    if scale is not None:
        clip_b32.logit_scale.data.fill_(math.log(scale))
        sd["logit_scale"] = clip_b32.logit_scale.detach().clone()
    try:
        model = _make_model(cuda, clip_b32)
        model.train()
        Bn = 8
        pix, ids, mask = O.synthetic_batch(Bn, seed=2)
        if vary_tok0:
            ids[:, 0] = torch.arange(Bn) * 37 + 5  # trainer.py:181's DummyDataset varies token 0
        pix, ids, mask = pix.to(cuda), ids.to(cuda), mask.to(cuda)
        ta, va = _adapters_sd(model)
        out = model(input_ids=ids, attention_mask=mask, pixel_values=pix, return_loss=True)
        out["loss"].backward()

        ta_r = {k: v.clone().requires_grad_(True) for k, v in ta.items()}
        va_r = {k: v.clone().requires_grad_(True) for k, v in va.items()}
        ref = O.model_m_forward(sd, 8, 12, ids, mask, pix, ta_r, va_r)
        ref["loss"].backward()
    finally:
        clip_b32.logit_scale.data.copy_(old)

    assert set(out.keys()) == set(ref.keys())
    loss_tol = LOSS_TOL if scale is None else 10 * LOSS_TOL  # a 7x larger scale magnifies the same feature error
    assert abs(out["loss"].item() - ref["loss"].item()) < loss_tol, (out["loss"].item(), ref["loss"].item())
    assert _rel(out["logits_per_text"], ref["logits_per_text"]) < LOGIT_TOL
    assert torch.allclose(out["logits_per_image"], out["logits_per_text"].t())
    assert _rel(out["image_features"], ref["image_features"]) < FEAT_TOL
    assert _rel(out["text_features"], ref["text_features"]) < FEAT_TOL
    if vary_tok0:
        assert torch.equal(out["logits_per_image"].argmax(1), ref["logits_per_image"].argmax(1))
    else:
        # reference quirk: every caption shares token 0, so all text rows are identical
        assert (out["text_features"] - out["text_features"][0]).abs().max().item() == 0.0
        assert abs(ref["loss"].item() - math.log(Bn)) < 5e-3
    # adapter-only gradients
    for name, mod, refd in (("text", model.text_adapter, ta_r), ("vision", model.vision_adapter, va_r)):
        for k, p in mod.named_parameters():
            g, gr = p.grad, refd[k].grad
            assert g is not None, (name, k)
            cos = torch.nn.functional.cosine_similarity(g.flatten(), gr.flatten(), dim=0).item()
            if scale is not None:
                denom = gr.abs().max().item() + 1e-12
                # scale 100 multiplies the backbone's bf16 feature error into the softmax: 5-6 % of the max element
                assert (g - gr).abs().max().item() / denom < 2e-1, (name, k, (g - gr).abs().max().item(), denom)
                assert cos > 0.99, (name, k, cos)
            elif vary_tok0 or name == "vision":
                assert cos > 0.98, (name, k, cos)
    assert all(p.grad is None for p in model.clip.parameters())


def test_loss_kernel_on_oracle_features(cuda, clip_b32, sd_b32):
    """fp32 loss within 1e-4 when the loss kernel sees the oracle's own features (isolates the bf16 backbone)."""
    from vlm_clip_b200 import ops

    pix, ids, mask = O.synthetic_batch(8, seed=2)
    ids[:, 0] = torch.arange(8) * 11 + 3
    pix, ids, mask = pix.to(cuda), ids.to(cuda), mask.to(cuda)
    with torch.no_grad():
        t = O.model_m_text_features(sd_b32, 8, ids, mask)
        i = O.model_m_image_features(sd_b32, 12, pix)
        ref = O.contrastive_loss(t, i, sd_b32["logit_scale"])
    loss, *_ = ops.clip_loss(t.contiguous(), i.contiguous(), float(sd_b32["logit_scale"].exp()))
    assert abs(loss.item() - ref["loss"].item()) < 1e-4


def test_trainer_step_matches_reference_step(cuda, clip_b32, sd_b32):
    """Three steps of CLIPAdapterTrainer.training_step against the oracle's forward + torch's clip_grad_norm_/AdamW
    (trainer.py:73-99) from the same initial adapters."""
    from vlm_clip_b200.trainer import CLIPAdapterTrainer

    model = _make_model(cuda, clip_b32, seed=5)
    ta, va = _adapters_sd(model)
    ref_params = {("t", k): torch.nn.Parameter(v.clone()) for k, v in ta.items()}
    ref_params.update({("v", k): torch.nn.Parameter(v.clone()) for k, v in va.items()})
    ref_opt = torch.optim.AdamW(list(ref_params.values()), lr=5e-5, weight_decay=0.01)
    trainer = CLIPAdapterTrainer(model, train_dataloader=[None] * 3, output_dir="/tmp/vlmclip_test_ckpt")
    for step in range(3):
        pix, ids, mask = O.synthetic_batch(8, seed=10 + step)
        ids[:, 0] = torch.randint(0, 1000, (8,), generator=torch.Generator().manual_seed(step))
        batch = {"input_ids": ids, "attention_mask": mask, "pixel_values": pix}
        loss = trainer.training_step(batch)
        ref_opt.zero_grad()
        tr = {k: ref_params[("t", k)] for k in ta}
        vr = {k: ref_params[("v", k)] for k in va}
        ref = O.model_m_forward(sd_b32, 8, 12, ids.to(cuda), mask.to(cuda), pix.to(cuda), tr, vr)
        ref["loss"].backward()
        torch.nn.utils.clip_grad_norm_(list(ref_params.values()), 1.0)
        ref_opt.step()
        assert abs(loss.item() - ref["loss"].item()) < LOSS_TOL
    for tag, mod, init in (("t", model.text_adapter, ta), ("v", model.vision_adapter, va)):
        for k, p in mod.named_parameters():
            # Adam normalises every element's step to ~lr whatever the gradient scale, so near-zero gradient elements
            # may step in either direction: compare the UPDATE vectors by direction and bound them by 3 steps * lr
            d_mine = (p.detach() - init[k]).flatten()
            d_ref = (ref_params[(tag, k)].detach() - init[k]).flatten()
            assert d_mine.abs().max().item() <= 3 * 5e-5 * 1.05 + 1e-7, k
            if d_ref.norm().item() > 0:
                cos = torch.nn.functional.cosine_similarity(d_mine, d_ref, dim=0).item()
                assert cos > 0.9, (tag, k, cos)


def test_model_errors_and_api(cuda, clip_b32):
    from vlm_clip_b200 import _native as N
    from vlm_clip_b200.model_m import CLIPWithAdapters

    pix, ids, mask = O.synthetic_batch(2)
    m_cpu = CLIPWithAdapters(clip=O.build_hf_clip(B32, seed=0, vision_layers=1, text_layers=1), use_shared_adapters=False)
    with pytest.raises(N.NativeError):  # CPU model / CPU tensors: no fallback
        m_cpu(input_ids=ids, attention_mask=mask, pixel_values=pix)
    m2 = _make_model(cuda, clip_b32)
    out = m2(input_ids=ids.to(cuda), attention_mask=mask.to(cuda), pixel_values=pix.to(cuda), return_loss=False)
    assert set(out) == {"text_features", "image_features"} and out["image_features"].shape == (2, 512)
    with pytest.raises(ValueError):
        m2.get_image_features(torch.zeros(1, 3, 128, 128, device=cuda))


def test_video_clips_and_uint8_frames(cuda, clip_b32, sd_b32):
    """SURVEY.md 8a-12 (config 4): clip feature = mean over frames of get_image_features; decoded uint8 frames give the
    same result as the float pixels the CPU preprocessing (oracle, pinned on cv2) produces from them."""
    import numpy as np

    from oracle import preprocess_oracle as P

    model = _make_model(cuda, clip_b32)
    model.eval()
    model.pixel_mean, model.pixel_std = P.IMAGENET_MEAN, P.IMAGENET_STD  # process_video.py:24
    rng = np.random.default_rng(5)
    B, T = 2, 3
    frames = rng.integers(0, 256, (B, T, 120, 160, 3), dtype=np.uint8)
    pix = P.preprocess_frames(frames.reshape(-1, 120, 160, 3), 224, 224, P.IMAGENET_MEAN, P.IMAGENET_STD)  # [B*T,3,224,224]
    clips = torch.from_numpy(pix).view(B, T, 3, 224, 224).permute(0, 2, 1, 3, 4).contiguous().to(cuda)  # [B,3,T,H,W]
    with torch.no_grad():
        per_frame = model.get_image_features(torch.from_numpy(pix).to(cuda))
        v_float = model.get_video_features(clips)
        v_u8 = model.get_video_features(torch.from_numpy(frames).to(cuda))
        f_u8 = model.get_image_features(torch.from_numpy(frames[0]).to(cuda))
    assert torch.allclose(v_float, per_frame.view(B, T, -1).mean(1), atol=1e-6)
    assert torch.equal(v_u8, v_float)            # same bf16 im2col bits -> identical features
    assert torch.equal(f_u8, per_frame[:T])
    # against the fp32 oracle of the per-frame path
    ta, va = _adapters_sd(model)
    ref = O.model_m_image_features(sd_b32, 12, torch.from_numpy(pix).to(cuda), va) if hasattr(O, "model_m_image_features") else None
    if ref is not None:
        assert _rel(v_float, ref.view(B, T, -1).mean(1)) < FEAT_TOL
    # forward() with clips: loss over clip-level features
    ids = torch.randint(3, 49406, (B, 77), device=cuda)
    ids[:, 0] = torch.tensor([5, 42], device=cuda)
    out = model(input_ids=ids, attention_mask=torch.ones_like(ids), pixel_values=torch.from_numpy(frames).to(cuda))
    assert out["logits_per_image"].shape == (B, B) and torch.isfinite(out["loss"])


def test_text_token0_shortcut_is_result_preserving(cuda, clip_b32):
    """Track M pools the causal text tower at token 0, so running the tower on token 0 alone gives the same features
    (SURVEY.md 8a-6 / 8d); the flag is opt-in and must not change any output."""
    model = _make_model(cuda, clip_b32)
    model.eval()
    pix, ids, mask = O.synthetic_batch(6, seed=4)
    ids[:, 0] = torch.arange(6) * 11 + 3
    mask[2, 9:] = 0
    pix, ids, mask = pix.to(cuda), ids.to(cuda), mask.to(cuda)
    with torch.no_grad():
        dense = model(input_ids=ids, attention_mask=mask, pixel_values=pix)
        model.text_token0_only = True
        short = model(input_ids=ids, attention_mask=mask, pixel_values=pix)
    assert _rel(short["text_features"], dense["text_features"]) < 1e-5
    assert abs(short["loss"].item() - dense["loss"].item()) < 1e-5
    assert torch.equal(short["logits_per_image"].argmax(1), dense["logits_per_image"].argmax(1))
    # image side: last vision layer for the CLS row only (single-query attention instead of the tile kernel)
    model.text_token0_only = False
    model.vision_cls_only_last_layer = True
    with torch.no_grad():
        cls = model(input_ids=ids, attention_mask=mask, pixel_values=pix)
    assert _rel(cls["image_features"], dense["image_features"]) < 5e-3
    assert abs(cls["loss"].item() - dense["loss"].item()) < 1e-3
    assert torch.equal(cls["logits_per_image"].argmax(1), dense["logits_per_image"].argmax(1))


def test_config3_vit_l14_dims_with_peclip_adapters(cuda):
    """BASELINE config 3 at reduced depth / batch: ViT-L/14 dimensions (S = 257, D = 1024 / 768, patch 14, 16 / 12 heads)
    with PE-CLIP TextualAdapters in the adapter slots, against the fp32 oracle; gradients flow to the adapters only."""
    L14 = "openai/clip-vit-large-patch14"
    clip = O.build_hf_clip(L14, seed=0, vision_layers=2, text_layers=2).to(cuda)
    for p_ in clip.parameters():
        p_.requires_grad_(False)
    sd = {k: v.detach() for k, v in clip.state_dict().items()}
    from vlm_clip_b200.model_m import CLIPWithAdapters

    torch.manual_seed(3)
    model = CLIPWithAdapters(clip=clip, use_shared_adapters=False, adapter_kind="peclip").to(cuda)
    model.train()
    Bn = 4
    pix, ids, mask = O.synthetic_batch(Bn, seed=6)
    ids[:, 0] = torch.arange(Bn) * 101 + 7
    pix, ids, mask = pix.to(cuda), ids.to(cuda), mask.to(cuda)
    out = model(input_ids=ids, attention_mask=mask, pixel_values=pix)
    out["loss"].backward()
    ta = {k: v.detach().clone().requires_grad_(True) for k, v in model.text_adapter.state_dict().items()}
    va = {k: v.detach().clone().requires_grad_(True) for k, v in model.vision_adapter.state_dict().items()}
    t_hid = O.text_tower(sd, ids, mask, 12)
    v_hid = O.vision_tower(sd, pix, 16)
    t_feat = O.peclip_textual_adapter(t_hid, ta)[:, 0] @ sd["text_projection.weight"].t()
    i_feat = O.peclip_textual_adapter(v_hid, va)[:, 0] @ sd["visual_projection.weight"].t()
    ref = O.contrastive_loss(t_feat, i_feat, sd["logit_scale"])
    ref["loss"].backward()
    assert out["image_features"].shape == (Bn, 768)
    assert _rel(out["image_features"], ref["image_features"]) < FEAT_TOL
    assert _rel(out["text_features"], ref["text_features"]) < FEAT_TOL
    assert abs(out["loss"].item() - ref["loss"].item()) < LOSS_TOL
    assert torch.equal(out["logits_per_image"].argmax(1), ref["logits_per_image"].argmax(1))
    for mod, refd in ((model.text_adapter, ta), (model.vision_adapter, va)):
        for k, p_ in mod.named_parameters():
            assert "adapter" not in k and p_.grad is not None  # names carry "adapter" through the attribute prefix
            cos = torch.nn.functional.cosine_similarity(p_.grad.flatten(), refd[k].grad.flatten(), dim=0).item()
            assert cos > 0.98, (k, cos)
    assert all("adapter" in n for n, p_ in model.named_parameters() if p_.requires_grad)


@pytest.mark.parametrize("cls,B,S,D,H", [("ContextAdapter", 3, 257, 1024, 16), ("SharedAdapter", 2, 50, 768, 12),
                                         ("ContextAdapter", 5, 197, 768, 12)])
def test_peclip_attention_adapters_forward_backward(cuda, cls, B, S, D, H):
    """Row 8a-8b: ContextAdapter / SharedAdapter (adapter/peclip.py:21-48) = LayerNorm(MHSA(x, x, x) + x), forward and the
    hand-written backward (all six parameter tensors and the input) against autograd over the fp32 oracle.  The module
    computes on the bf16 tensor cores like the towers, so the bounds are bf16 ones (measured 3-6e-3 forward, 1-2e-2 on the
    gradients; same kernels and bounds as the full-fine-tune backward, tests/test_gpu_backward.py)."""
    from vlm_clip_b200.adapter import peclip

    torch.manual_seed(17)
    mod = getattr(peclip, cls)(D, H).to(cuda).train()
    with torch.no_grad():  # nn.MultiheadAttention / LayerNorm start with zero biases and unit gains: make them matter
        for p_ in (mod.mhsa.in_proj_bias, mod.mhsa.out_proj.bias, mod.layer_norm.bias):
            p_.normal_(0, 0.1)
        mod.layer_norm.weight.uniform_(0.5, 1.5)
    g = torch.Generator().manual_seed(18)
    x = (torch.randn(B, S, D, generator=g) * 0.7).to(cuda).requires_grad_(True)
    w = torch.randn(B, S, D, generator=g).to(cuda)
    y = mod(x)
    (y * w).sum().backward()
    a = {k: v.detach().clone().requires_grad_(True) for k, v in mod.state_dict().items()}
    xr = x.detach().clone().requires_grad_(True)
    ref = O.mhsa_adapter(xr, a, H)
    (ref * w).sum().backward()
    assert y.shape == ref.shape and y.dtype == torch.float32
    assert _rel(y, ref) < 1e-2, _rel(y, ref)
    assert _rel(x.grad, xr.grad) < 3e-2, _rel(x.grad, xr.grad)
    for k, p_ in mod.named_parameters():
        gr = a[k].grad
        assert p_.grad is not None and p_.grad.shape == gr.shape, k
        cos = torch.nn.functional.cosine_similarity(p_.grad.flatten(), gr.flatten(), dim=0).item()
        assert cos > 0.995 and _rel(p_.grad, gr) < 5e-2, (k, cos, _rel(p_.grad, gr))
    # only token 0 consumed (how Track M uses the adapter slot): gradient flows from that row alone
    mod.zero_grad(set_to_none=True)
    y0 = mod(x.detach())[:, 0, :]
    (y0 * w[:, 0]).sum().backward()
    for v in a.values():
        v.grad = None
    (O.mhsa_adapter(x.detach(), a, H)[:, 0, :] * w[:, 0]).sum().backward()
    for k, p_ in mod.named_parameters():
        cos = torch.nn.functional.cosine_similarity(p_.grad.flatten(), a[k].grad.flatten(), dim=0).item()
        assert cos > 0.99, (k, cos)


def test_track_m_with_context_adapter_in_the_vision_slot(cuda):
    """adapter_kind="peclip_context" (the option BASELINE config 3's "PE-CLIP adapter" can use): TextualAdapter on the
    text side, ContextAdapter over all 257 patch tokens on the vision side, trained through the loss; eager and the
    two-graph replay agree bit for bit."""
    from vlm_clip_b200.model_m import CLIPWithAdapters
    from vlm_clip_b200.trainer import CLIPAdapterTrainer

    L14 = "openai/clip-vit-large-patch14"
    clip = O.build_hf_clip(L14, seed=0, vision_layers=2, text_layers=2).to(cuda)
    for p_ in clip.parameters():
        p_.requires_grad_(False)
    sd = {k: v.detach() for k, v in clip.state_dict().items()}
    torch.manual_seed(3)
    model = CLIPWithAdapters(clip=clip, use_shared_adapters=False, adapter_kind="peclip_context").to(cuda).train()
    Bn = 4
    pix, ids, mask = O.synthetic_batch(Bn, seed=6)
    ids[:, 0] = torch.arange(Bn) * 101 + 7
    pix, ids, mask = pix.to(cuda), ids.to(cuda), mask.to(cuda)
    out = model(input_ids=ids, attention_mask=mask, pixel_values=pix)
    out["loss"].backward()
    ta = {k: v.detach().clone().requires_grad_(True) for k, v in model.text_adapter.state_dict().items()}
    va = {k: v.detach().clone().requires_grad_(True) for k, v in model.vision_adapter.state_dict().items()}
    t_feat = O.peclip_textual_adapter(O.text_tower(sd, ids, mask, 12), ta)[:, 0] @ sd["text_projection.weight"].t()
    i_feat = O.mhsa_adapter(O.vision_tower(sd, pix, 16), va, 16)[:, 0] @ sd["visual_projection.weight"].t()
    ref = O.contrastive_loss(t_feat, i_feat, sd["logit_scale"])
    ref["loss"].backward()
    assert _rel(out["image_features"], ref["image_features"]) < FEAT_TOL
    assert abs(out["loss"].item() - ref["loss"].item()) < 2 * LOSS_TOL
    for k, p_ in model.vision_adapter.named_parameters():
        cos = torch.nn.functional.cosine_similarity(p_.grad.flatten(), va[k].grad.flatten(), dim=0).item()
        assert cos > 0.97, (k, cos)
    assert all(p_.grad is None for p_ in clip.parameters())
    batch = {"input_ids": ids, "attention_mask": mask, "pixel_values": pix}
    losses = {}
    for graph in (False, True):
        torch.manual_seed(3)
        m = CLIPWithAdapters(clip=clip, use_shared_adapters=False, adapter_kind="peclip_context").to(cuda).train()
        tr = CLIPAdapterTrainer(m, [None], learning_rate=1e-3, output_dir="/tmp/vlmclip_ctx", cuda_graph=graph, graph_warmup_steps=2)
        losses[graph] = torch.stack([tr.training_step(batch).clone() for _ in range(5)])
    assert torch.equal(losses[True], losses[False]) and losses[False][-1] < losses[False][0]


def test_shared_mhs_adapter_inference(cuda, clip_b32):
    """Row 8a-8: the cross-modal adapter's inference path on the GPU against the fp32 oracle, stand-alone and inside
    CLIPWithAdapters (token-0 evaluation, model_m.py:93-102)."""
    from vlm_clip_b200 import _native as N
    from vlm_clip_b200.adapter.clip_adapter import SharedMHSAttentionAdapter
    from vlm_clip_b200.model_m import CLIPWithAdapters

    torch.manual_seed(11)
    mod = SharedMHSAttentionAdapter().to(cuda).eval()
    g = torch.Generator().manual_seed(12)
    xt = torch.randn(3, 77, 512, generator=g).to(cuda)
    table = (torch.randn(1, 50, 768, generator=g) * 0.5).to(cuda)
    a = {k: v.detach() for k, v in mod.state_dict().items()}
    with torch.no_grad():
        y = mod(xt, table)
    ref = O.shared_mhs_adapter(xt, table, a)
    assert y.shape == ref.shape and _rel(y, ref) < 1e-2, _rel(y, ref)

    torch.manual_seed(2)
    model = CLIPWithAdapters(clip=clip_b32, use_shared_adapters=True, shared_adapter_layers=2).to(cuda).eval()
    pix, ids, mask = O.synthetic_batch(4, seed=8)
    ids[:, 0] = torch.arange(4) * 13 + 2
    with torch.no_grad():
        t = model.get_text_features(ids.to(cuda), mask.to(cuda))
    sd = {k: v.detach() for k, v in clip_b32.state_dict().items()}
    hid = O.seq_adapter(O.text_tower(sd, ids.to(cuda), mask.to(cuda), 8),
                        {k: v.detach() for k, v in model.text_adapter.state_dict().items()})
    tab = sd["vision_model.embeddings.position_embedding.weight"].unsqueeze(0)
    for ad in model.shared_adapters:
        hid = O.shared_mhs_adapter(hid, tab, {k: v.detach() for k, v in ad.state_dict().items()})
    ref_t = hid[:, 0] @ sd["text_projection.weight"].t()
    assert _rel(t, ref_t) < FEAT_TOL, _rel(t, ref_t)


def test_shared_mhs_adapter_training(cuda, clip_b32):
    """Row 8a-8, trainable path: output and the gradients of all 18 parameters, of the text rows and of the table against
    autograd over the fp32 oracle (dropout 0: the reference's dropout is stochastic), then a training-mode run with
    dropout 0.1, then Track M end to end with two shared adapter layers in the loss."""
    from vlm_clip_b200.adapter.clip_adapter import SharedMHSAttentionAdapter
    from vlm_clip_b200.model_m import CLIPWithAdapters

    torch.manual_seed(21)
    mod = SharedMHSAttentionAdapter(dropout=0.0).to(cuda).train()
    g = torch.Generator().manual_seed(22)
    xt = torch.randn(6, 1, 512, generator=g).to(cuda).requires_grad_(True)
    table = (torch.randn(1, 50, 768, generator=g) * 0.5).to(cuda).requires_grad_(True)
    w = torch.randn(6, 1, 512, generator=g).to(cuda)
    y = mod(xt, table)
    (y * w).sum().backward()
    a = {k: v.detach().clone().requires_grad_(True) for k, v in mod.state_dict().items()}
    xr, tr_ = xt.detach().clone().requires_grad_(True), table.detach().clone().requires_grad_(True)
    ref = O.shared_mhs_adapter(xr, tr_, a)
    (ref * w).sum().backward()
    assert _rel(y, ref) < 1e-5, _rel(y, ref)          # fp32 path
    assert _rel(xt.grad, xr.grad) < 1e-4 and _rel(table.grad, tr_.grad) < 1e-4
    for k, p_ in mod.named_parameters():
        assert p_.grad is not None and _rel(p_.grad, a[k].grad) < 1e-4, (k, _rel(p_.grad, a[k].grad))

    # dropout 0.1 (adapter/clip_adapter.py:84,96): stochastic in train mode, reproducible under a seed, off in eval mode
    torch.manual_seed(31)
    drop = SharedMHSAttentionAdapter().to(cuda).train()
    torch.manual_seed(5)
    y1 = drop(xt.detach(), table.detach())
    torch.manual_seed(5)
    y2 = drop(xt.detach(), table.detach())
    y3 = drop(xt.detach(), table.detach())
    assert torch.equal(y1, y2) and not torch.equal(y1, y3)
    drop.eval()
    with torch.no_grad():
        ye = drop(xt.detach(), table.detach())  # bf16 tensor-core inference path
    ref_e = O.shared_mhs_adapter(xt.detach(), table.detach(), {k: v.detach() for k, v in drop.state_dict().items()})
    assert _rel(ye, ref_e) < 1e-2
    assert 1e-3 < _rel(y1, ref_e) < 0.5  # dropped, but the same function in expectation

    # Track M with the reference's default constructor flags (use_shared_adapters=True, two layers), batch > 1
    torch.manual_seed(2)
    model = CLIPWithAdapters(clip=clip_b32, use_shared_adapters=True, shared_adapter_layers=2).to(cuda).train()
    for ad in model.shared_adapters:
        ad.cross_attn.dropout = 0.0
        ad.mlp[3].p = 0.0
    pix, ids, mask = O.synthetic_batch(4, seed=8)
    ids[:, 0] = torch.arange(4) * 13 + 2
    pix, ids, mask = pix.to(cuda), ids.to(cuda), mask.to(cuda)
    out = model(input_ids=ids, attention_mask=mask, pixel_values=pix, return_loss=True)
    out["loss"].backward()
    sd = {k: v.detach() for k, v in clip_b32.state_dict().items()}
    ta = {k: v.detach().clone().requires_grad_(True) for k, v in model.text_adapter.state_dict().items()}
    va = {k: v.detach().clone().requires_grad_(True) for k, v in model.vision_adapter.state_dict().items()}
    sa = [{k: v.detach().clone().requires_grad_(True) for k, v in ad.state_dict().items()} for ad in model.shared_adapters]
    hid = O.seq_adapter(O.text_tower(sd, ids, mask, 8), ta)
    tab = sd["vision_model.embeddings.position_embedding.weight"].unsqueeze(0)
    for a_ in sa:
        hid = O.shared_mhs_adapter(hid, tab, a_)
    ref_t = hid[:, 0] @ sd["text_projection.weight"].t()
    ref_i = O.model_m_image_features(sd, 12, pix, va)
    ref_out = O.contrastive_loss(ref_t, ref_i, sd["logit_scale"])
    ref_out["loss"].backward()
    # two shared-adapter layers (LayerNorm -> cross-attention -> MLP) sit between the bf16 text tower and the loss and
    # amplify its 6e-3 feature error; on 4 pairs the loss error is a noisy quantity (6.7e-4 and 1.1e-3 measured for two
    # builds whose LayerNorm statistics differ in the last fp32 bit), against 1.5e-4 without the shared adapters
    assert abs(out["loss"].item() - ref_out["loss"].item()) < 4 * LOSS_TOL
    assert _rel(out["text_features"], ref_out["text_features"]) < FEAT_TOL
    for ad, a_ in zip(model.shared_adapters, sa):
        for k, p_ in ad.named_parameters():
            gr = a_[k].grad
            assert p_.grad is not None, k
            if gr.norm().item() < 1e-12:
                continue
            cos = torch.nn.functional.cosine_similarity(p_.grad.flatten(), gr.flatten(), dim=0).item()
            assert cos > 0.97, (k, cos)
    assert all(p_.grad is None for p_ in model.clip.parameters())


def test_eval_after_optimizer_steps_sees_updated_weights(cuda, clip_b32):
    """the bf16 weight packs of the inference paths were keyed on Parameter._version only, which
    FusedAdamW (raw-pointer updates of the arena) never moves, so `evaluate()` after an epoch ran the first epoch's
    matrices.  eval -> N optimiser steps -> eval must follow the fp32 trainable path evaluated on the new weights."""
    from vlm_clip_b200.model_m import CLIPWithAdapters
    from vlm_clip_b200.trainer import CLIPAdapterTrainer

    torch.manual_seed(2)
    model = CLIPWithAdapters(clip=clip_b32, use_shared_adapters=True, shared_adapter_layers=1).to(cuda)
    for ad in model.shared_adapters:  # deterministic trainable path for the comparison
        ad.cross_attn.dropout = 0.0
        ad.mlp[3].p = 0.0
    pix, ids, mask = O.synthetic_batch(4, seed=8)
    ids[:, 0] = torch.arange(4) * 13 + 2
    pix, ids, mask = pix.to(cuda), ids.to(cuda), mask.to(cuda)
    model.eval()
    with torch.no_grad():
        t_before = model.get_text_features(ids, mask).clone()  # packs the bf16 copies
    trainer = CLIPAdapterTrainer(model, [None], learning_rate=3e-3, output_dir="/tmp/vlmclip_test_stale")
    model.train()
    for _ in range(4):
        trainer.training_step({"input_ids": ids, "attention_mask": mask, "pixel_values": pix})
    model.eval()
    with torch.no_grad():
        t_after = model.get_text_features(ids, mask)     # bf16 inference path, must re-pack
    model.train()
    t_ref = model.get_text_features(ids, mask).detach()  # fp32 trainable path on the live parameters
    assert _rel(t_before, t_ref) > 5e-2                  # the steps moved the function ...
    assert _rel(t_after, t_ref) < 2e-2, (_rel(t_after, t_ref), _rel(t_before, t_ref))  # ... and eval follows it


def test_cuda_graph_step_matches_eager_step(cuda, clip_b32, tmp_path):
    """CLIPAdapterTrainer(cuda_graph=True): the captured-and-replayed step must be the eager step, bit for bit (same
    kernels, same order, deterministic reductions), including the learning-rate schedule that is pushed to the device
    outside the graph and a change of batch shape (new signature -> eager warm-up -> second graph)."""
    from vlm_clip_b200.trainer import CLIPAdapterTrainer

    def batch(seed, n):
        pix, ids, mask = O.synthetic_batch(n, seed=seed)
        ids[:, 0] = (torch.arange(n) * 7 + seed) % 1000
        return {"input_ids": ids.to(cuda), "attention_mask": mask.to(cuda), "pixel_values": pix.to(cuda)}

    seq = [batch(30 + i, 4) for i in range(7)] + [batch(50 + i, 6) for i in range(5)]
    res = {}
    for mode in (False, True):
        m = _make_model(cuda, clip_b32, seed=9)
        tr = CLIPAdapterTrainer(m, seq, learning_rate=1e-3, warmup_steps=3, output_dir=str(tmp_path / f"g{int(mode)}"),
                                cuda_graph=mode, graph_warmup_steps=2)
        tr._total_steps = len(seq)
        tr.optimizer.set_lr(0.0)
        m.train()
        losses = [tr.training_step(b).clone() for b in seq]
        torch.cuda.synchronize()
        res[mode] = (torch.stack(losses), tr.optimizer.flat.clone(), tr.optimizer.exp_avg_sq.clone(), tr)
    tr_g = res[True][3]
    assert tr_g.graph_replays == (7 - 2) + (5 - 2) and len(tr_g._graphs) == 2 and tr_g.graph_launches_per_step > 50
    assert res[False][3].graph_replays == 0
    assert torch.equal(res[True][0], res[False][0]), (res[True][0] - res[False][0]).abs().max()
    assert torch.equal(res[True][1], res[False][1]) and torch.equal(res[True][2], res[False][2])
    assert int(tr_g.optimizer.step_t.item()) == len(seq)
    # evaluation / no-grad forward after graphed training sees the updated adapters (they live in the arena the graph writes)
    tr_g.model.eval()
    with torch.no_grad():
        out_g = tr_g.model(**seq[0])
        res[False][3].model.eval()
        out_e = res[False][3].model(**seq[0])
    assert torch.equal(out_g["loss"], out_e["loss"])


def test_reference_trainer_loop_drives_the_mirror_model(cuda, clip_b32, tmp_path):
    """The drop-in claim at the trainer boundary: the reference's OWN, unmodified
    `trainer.CLIPAdapterTrainer.train` loop (trainer.py:50-125: name filter, torch.optim.AdamW, linear warm-up schedule,
    clip_grad_norm_, loss.item()) runs over the mirror `CLIPWithAdapters` bound under the reference's module name, as
    INTEGRATION.md describes, and takes the same trajectory as the mirror trainer (fused clip + AdamW) on the same
    batches.  Needs the reference's files staged by oracle/stage_reference.py (they travel with the snapshot)."""
    import importlib
    import sys

    from oracle import ref_harness as H

    if not H.available():
        pytest.skip("oracle/_ref/reference is not staged (run oracle/stage_reference.py where /root/reference exists)")
    import vlm_clip_b200.model_m as mirror_model_m
    from vlm_clip_b200.trainer import CLIPAdapterTrainer as MirrorTrainer

    saved = {k: sys.modules.get(k) for k in ("model_m", "trainer")}
    sys.modules["model_m"] = mirror_model_m            # the binding INTEGRATION.md prescribes
    sys.modules.pop("trainer", None)
    sys.path.insert(0, str(H.STAGED))
    try:
        ref_trainer = importlib.import_module("trainer")  # the reference's trainer.py, byte for byte
        assert ref_trainer.CLIPWithAdapters is mirror_model_m.CLIPWithAdapters

        def batches():
            out = []
            for s_ in range(4):
                pix, ids, mask = O.synthetic_batch(4, seed=60 + s_)
                ids[:, 0] = torch.arange(4) * 9 + s_
                out.append({"input_ids": ids, "attention_mask": mask, "pixel_values": pix})
            return out

        m_ref = _make_model(cuda, clip_b32, seed=11)
        tr_ref = ref_trainer.CLIPAdapterTrainer(m_ref, batches(), learning_rate=1e-3, warmup_steps=2,
                                                output_dir=str(tmp_path / "ref"))
        assert len(tr_ref.trainable_params) == 12 and isinstance(tr_ref.optimizer, torch.optim.AdamW)
        tr_ref.train(num_epochs=2)                      # 8 steps through the reference's loop
        m_mir = _make_model(cuda, clip_b32, seed=11)
        tr_mir = MirrorTrainer(m_mir, batches(), learning_rate=1e-3, warmup_steps=2, output_dir=str(tmp_path / "mir"),
                               log_every=1000)
        tr_mir.train(num_epochs=2)
    finally:
        sys.path.remove(str(H.STAGED))
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    for (n1, p1), (n2, p2) in zip(m_ref.named_parameters(), m_mir.named_parameters()):
        if "adapter" not in n1:
            continue
        assert n1 == n2
        step = (p1 - p2).abs().max().item()
        # 8 Adam steps of ~1e-3 each: the two optimisers (torch's foreach AdamW vs the fused kernel) agree to fp32 rounding
        assert step < 2e-5, (n1, step)
    assert (tmp_path / "ref" / "final_adapter.pt").exists()  # the reference's save path works on the mirror's checkpoint API
    blob = torch.load(tmp_path / "ref" / "final_adapter.pt")
    assert set(blob) == {"text_adapter", "vision_adapter"}


def test_resume_continues_like_an_uninterrupted_run(cuda, clip_b32, tmp_path):
    """SURVEY.md 8f-4 (the reference has no resume, trainer.py:157-167): 6 steps in one go == 3 steps, save, fresh
    trainer + model, load, 3 more steps.  Compares parameters, both Adam moments, the step counter and the position in
    the warm-up / decay schedule."""
    from vlm_clip_b200.trainer import CLIPAdapterTrainer

    def batches():
        out = []
        for s in range(3):
            pix, ids, mask = O.synthetic_batch(4, seed=20 + s)
            ids[:, 0] = torch.arange(4) * 7 + s
            out.append({"input_ids": ids, "attention_mask": mask, "pixel_values": pix})
        return out

    def fresh():
        m = _make_model(cuda, clip_b32, seed=9)
        return m, CLIPAdapterTrainer(m, batches(), learning_rate=1e-3, warmup_steps=2, output_dir=str(tmp_path / "ck"),
                                     log_every=1000)

    m_a, tr_a = fresh()
    tr_a.train(num_epochs=2, save_every=100, eval_every=100)      # 6 steps, uninterrupted
    m_b, tr_b = fresh()                                            # the interrupted run: 3 steps of the 2-epoch schedule
    tr_b._total_steps = 6
    tr_b.optimizer.set_lr(0.0)                                     # step 0 of a 2-step warm-up
    m_b.train()
    for b in batches():
        tr_b.training_step(b)
    state = str(tmp_path / "state.pt")
    tr_b.save_training_state(state)
    assert tr_b._global_step == 3
    m_c, tr_c = fresh()                                            # new process: fresh model + trainer
    for p_ in tr_c.trainable_params:
        p_.data.add_(1.0)                                          # make sure the parameters come from the file
    tr_c.load_training_state(state)
    tr_c.train(num_epochs=2, save_every=100, eval_every=100)      # skips the 3 completed steps, runs 3
    assert tr_c._global_step == 6 == tr_a._global_step
    assert int(tr_c.optimizer.step_t.item()) == 6
    for name in ("flat", "exp_avg", "exp_avg_sq"):
        a, c = getattr(tr_a.optimizer, name), getattr(tr_c.optimizer, name)
        assert torch.equal(a, c), (name, (a - c).abs().max().item())
    assert tr_a.optimizer.param_groups[0]["lr"] == tr_c.optimizer.param_groups[0]["lr"]


def test_full_size_properties_vit_b16_batch256(cuda):
    """BASELINE config 2 at full size (ViT-B/16, 256 pairs), checked through size-independent properties instead of the
    oracle (a fp32 CPU pass at this size takes minutes): (1) every caption starts with BOS and the text tower is causal,
    so all text rows coincide and the symmetric InfoNCE loss is ln(256); (2) the step is deterministic
    (no atomics, fixed reduction orders): two runs are bit identical; (3) permuting the batch permutes the image
    features; (4) only adapter parameters receive gradients and they are finite."""
    B16 = "openai/clip-vit-base-patch16"
    clip = O.build_hf_clip(B16, seed=0).to(cuda)
    for p_ in clip.parameters():
        p_.requires_grad_(False)
    model = _make_model(cuda, clip)
    model.train()
    Bn = 256
    pix, ids, mask = O.synthetic_batch(Bn, seed=2)
    pix, ids, mask = pix.to(cuda), ids.to(cuda), mask.to(cuda)
    out1 = model(input_ids=ids, attention_mask=mask, pixel_values=pix)
    out1["loss"].backward()
    g1 = [p_.grad.clone() for p_ in model.parameters() if p_.requires_grad]
    assert abs(out1["loss"].item() - math.log(Bn)) < 2e-3, out1["loss"].item()
    assert (out1["text_features"] - out1["text_features"][0]).abs().max().item() == 0.0
    assert all(torch.isfinite(g).all() for g in g1) and len(g1) == 12
    assert all(p_.grad is None for p_ in clip.parameters())
    model.zero_grad(set_to_none=True)
    out2 = model(input_ids=ids, attention_mask=mask, pixel_values=pix)
    out2["loss"].backward()
    g2 = [p_.grad.clone() for p_ in model.parameters() if p_.requires_grad]
    assert torch.equal(out1["image_features"], out2["image_features"]) and torch.equal(out1["loss"], out2["loss"])
    assert all(torch.equal(a, b) for a, b in zip(g1, g2))
    perm = torch.randperm(Bn, generator=torch.Generator().manual_seed(0)).to(cuda)
    with torch.no_grad():
        f = model.get_image_features(pix)
        fp = model.get_image_features(pix[perm].contiguous())
    assert torch.equal(fp, f[perm])
