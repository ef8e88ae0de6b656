"""Full fine-tune execution of the CLIP towers (BASELINE config 5; reference: `CLIPWithAdapters(freeze_clip=False)`,
model_m.py:22,72-75, where autograd differentiates HF modeling_clip.py).

`TrainableClipTowers` runs the towers straight from the LIVE fp32 parameters of the HuggingFace `CLIPModel` (no packed
copy, no LayerNorm folding: the weights change every step) and implements the backward by hand, one
`torch.autograd.Function` per tower, so `loss.backward()` fills `.grad` of every CLIP parameter the loss depends on.

Forward per layer (HF:363-384), every intermediate the backward needs is kept in bf16:

    xn1 = LN1(x) -> qkv = xn1 Wqkv^T + b -> att = attention(qkv) -> x_mid = x + att Wo^T + b
    xn2 = LN2(x_mid) -> a = xn2 W1^T + b -> hid = quick_gelu(a) -> x_out = x_mid + hid W2^T + b

Backward: every dense product goes through the SAME tcgen05 GEMM kernel as the forward, C = A W^T, by transposing
operands (`vlmclip_transpose_to_bf16`):  dX = dY (W^T)^T  [A = dY, W = W^T],  dW = dY^T X = (dY^T)(X^T)^T  [A = dY^T,
W = X^T, fp32 output],  db = row sums of dY^T.  The gradient of the residual stream is carried in fp32 (with a bf16
copy as GEMM operand), LayerNorm / quick_gelu / attention backward are the kernels of csrc/backward.cu and
csrc/attention_bwd.cu.  What is known to be slow and planned next: the weight-gradient GEMMs have few output tiles
(no split-K yet) and the attention backward is fp32 SIMT.

Track M pools both towers at token 0 (model_m.py:102,122), so the tower functions return those rows only: the text
one after final_layer_norm (HF:562), the vision one without post_layernorm (never applied by the reference).
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import _native as N
from . import ops

bf16, f32 = torch.bfloat16, torch.float32
import os as _os

# weight-gradient GEMMs: "mn" (default) MN-major operands + split reduction, "splitk" transposed copies + split
# reduction, "plain" transposed copies, one reduction per output tile (round 1); A/B switch
_WGRAD = _os.environ.get("VLMCLIP_WGRAD", "mn")

_LAYER_PARAMS = ("layer_norm1.weight", "layer_norm1.bias", "self_attn.q_proj.weight", "self_attn.q_proj.bias",
                 "self_attn.k_proj.weight", "self_attn.k_proj.bias", "self_attn.v_proj.weight", "self_attn.v_proj.bias",
                 "self_attn.out_proj.weight", "self_attn.out_proj.bias", "layer_norm2.weight", "layer_norm2.bias",
                 "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias")
_NLP = len(_LAYER_PARAMS)


def _dense_bwd(dy16, x16, W16):
    """Backward of y = x W^T + b through the forward GEMM.  dy16 [M, N], x16 [M, K], W16 [N, K] (all bf16) ->
    (dx bf16 [M, K], dW fp32 [N, K], db fp32 [N])."""
    dx = ops.gemm(dy16, ops.transpose_bf16(W16))
    if _WGRAD == "mn":
        # dW = dY^T X straight from dY and X (MN-major operand descriptors), the token reduction split over the SMs;
        # db = column sums of dY.  No transposed copies of the two activation-sized tensors.
        return dx, ops.gemm_atb_splitk(dy16, x16), ops.colsum_bf16(dy16)
    dyT = ops.transpose_bf16(dy16)
    # few output tiles (N x K of the layer), a 50 k-row reduction: the reduction is what gets spread over the SMs
    xT = ops.transpose_bf16(x16)
    dW = ops.gemm_splitk(dyT, xT) if _WGRAD == "splitk" else ops.gemm(dyT, xT, out_fp32=True)
    return dx, dW, ops.rowsum_bf16(dyT)


def _encoder_fwd(x, lp: List[torch.Tensor], n_layers: int, B: int, S: int, H: int, eps: float, causal: bool, key_mask,
                 keep: bool):
    """All encoder layers from fp32 parameters `lp` (16 per layer, _LAYER_PARAMS order).  Returns (x_out, saved)."""
    saved = []
    D = x.shape[1]
    for l in range(n_layers):
        (g1, b1, qw, qb, kw, kb, vw, vb, ow, ob, g2, b2, f1w, f1b, f2w, f2b) = lp[l * _NLP:(l + 1) * _NLP]
        Wqkv = torch.empty((3 * D, D), device=x.device, dtype=bf16)
        for j, w in enumerate((qw, kw, vw)):
            ops.cast_bf16(w, out=Wqkv[j * D:(j + 1) * D])
        bqkv = torch.cat([qb, kb, vb])  # 3D floats: a copy, no arithmetic
        Wo, W1, W2 = ops.cast_bf16(ow), ops.cast_bf16(f1w), ops.cast_bf16(f2w)
        xn1 = ops.layernorm(x, g1, b1, eps)
        qkv = ops.gemm(xn1, Wqkv, bias=bqkv)
        att = ops.attention(qkv, B, S, H, causal=causal, key_mask=key_mask)
        x_mid = ops.gemm(att, Wo, bias=ob, residual=x)
        xn2 = ops.layernorm(x_mid, g2, b2, eps)
        a = ops.gemm(xn2, W1, bias=f1b)
        hid = ops.quick_gelu(a)
        x_out = ops.gemm(hid, W2, bias=f2b, residual=x_mid)
        if keep:
            saved.append((x, xn1, qkv, att, x_mid, xn2, a, Wqkv, Wo, W1, W2))
        x = x_out
    return x, saved


def _encoder_bwd(dx32, saved, lp, B: int, S: int, H: int, eps: float, causal: bool, key_mask, sink=None):
    """Reverse pass over the layers.  dx32: fp32 gradient of the encoder output [M, D].  Returns (dx32 of the encoder
    input, list of parameter gradients in `lp` order).

    sink(params, grads) (optional): called as soon as a layer's 16 parameter gradients exist, with the layer's
    parameters; it takes the gradients over (writes them into `.grad` and may start their all-reduce while the reverse
    pass continues with the layers below) and this function then reports None for them, so autograd does not
    accumulate them a second time."""
    D = dx32.shape[1]
    grads: List[Optional[torch.Tensor]] = [None] * len(lp)
    dx16 = ops.cast_bf16(dx32)
    other32 = torch.empty_like(dx32)
    for l in range(len(saved) - 1, -1, -1):
        x, xn1, qkv, att, x_mid, xn2, a, Wqkv, Wo, W1, W2 = saved[l]
        g1, g2 = lp[l * _NLP + 0], lp[l * _NLP + 10]
        o = l * _NLP
        # ---- MLP: x_out = x_mid + fc2(quick_gelu(fc1(LN2(x_mid)))) ----
        hid = ops.quick_gelu(a)
        d_hid, grads[o + 14], grads[o + 15] = _dense_bwd(dx16, hid, W2)
        del hid
        d_a = ops.quick_gelu_bwd(a, d_hid)
        d_xn2, grads[o + 12], grads[o + 13] = _dense_bwd(d_a, xn2, W1)
        del d_a, d_hid
        dmid16 = torch.empty_like(dx16)
        grads[o + 10], grads[o + 11] = ops.layernorm_bwd(d_xn2, x_mid, g2, eps, dres=dx32, dx_f32=other32, dx_bf16=dmid16)
        dx32, other32 = other32, dx32  # dx32 now holds d x_mid
        # ---- attention: x_mid = x + out_proj(attention(qkv(LN1(x)))) ----
        d_att, grads[o + 8], grads[o + 9] = _dense_bwd(dmid16, att, Wo)
        d_qkv = ops.attention_bwd(qkv, att, d_att, B, S, H, causal=causal, key_mask=key_mask)
        d_xn1, dWqkv, dbqkv = _dense_bwd(d_qkv, xn1, Wqkv)
        for j in range(3):
            grads[o + 2 + 2 * j] = dWqkv[j * D:(j + 1) * D]
            grads[o + 3 + 2 * j] = dbqkv[j * D:(j + 1) * D]
        grads[o + 0], grads[o + 1] = ops.layernorm_bwd(d_xn1, x, g1, eps, dres=dx32, dx_f32=other32, dx_bf16=dx16)
        dx32, other32 = other32, dx32  # dx32 now holds d x (the layer's input)
        saved[l] = None  # release the layer's activations as the reverse pass moves on
        if sink is not None:
            sink(lp[o:o + _NLP], grads[o:o + _NLP])
            grads[o:o + _NLP] = [None] * _NLP
    return dx32, grads


class _VisionTowerFn(torch.autograd.Function):
    """pixel_values -> last_hidden_state[:, 0] as fp32 [B, D] (HF:667-691 without post_layernorm; model_m.py:110-122)."""

    @staticmethod
    def forward(ctx, tw, pixel_values, *params):
        patch_w, cls, pos, pre_g, pre_b = params[:5]
        lp = list(params[5:])
        keep = any(ctx.needs_input_grad[2:])
        B = pixel_values.shape[0]
        S, D, p = tw.Sv, tw.Dv, tw.patch
        cols = ops.im2col(pixel_values, p)  # bf16 [B*np, Kpad]
        K, Kpad = 3 * p * p, cols.shape[1]
        Wp = ops.cast_bf16(patch_w.reshape(D, K))
        if Kpad != K:  # patch 14: zero-padded K (a copy of the cast weight, no arithmetic)
            Wpp = torch.zeros((D, Kpad), device=Wp.device, dtype=bf16)
            Wpp[:, :K].copy_(Wp)
            Wp = Wpp
        patches = ops.gemm(cols, Wp)
        e = ops.vision_embed(patches, cls, pos, B, S)
        x0 = ops.layernorm(e, pre_g, pre_b, tw.eps_v)
        x, saved = _encoder_fwd(x0, lp, tw.Lv, B, S, tw.Hv, tw.eps_v, False, None, keep)
        out = ops.gather_rows_f32(x, B, S * D, D)
        if keep:
            ctx.tw, ctx.saved, ctx.lp, ctx.emb = tw, saved, lp, (cols, e, pre_g, B, K, Kpad, tuple(patch_w.shape))
        return out

    @staticmethod
    def backward(ctx, d_rows):
        tw = ctx.tw
        cols, e, pre_g, B, K, Kpad, pw_shape = ctx.emb
        S, D = tw.Sv, tw.Dv
        dx32 = torch.zeros((B * S, D), device=d_rows.device, dtype=f32)
        dx32.view(B, S, D)[:, 0].copy_(d_rows)  # only the CLS rows carry gradient
        dx32, lgrads = _encoder_bwd(dx32, ctx.saved, ctx.lp, B, S, tw.Hv, tw.eps_v, False, None, tw.layer_grad_sink)
        ctx.saved = None
        de32 = torch.empty_like(dx32)
        de16 = torch.empty((B * S, D), device=dx32.device, dtype=bf16)
        d_pre_g, d_pre_b = ops.layernorm_bwd(dx32, e, pre_g, tw.eps_v, dx_f32=de32, dx_bf16=de16)
        d_pos = ops.colsum_f32(de32.view(B, S * D)).view(S, D)
        d_cls = d_pos[0].clone()  # e[b,0] = cls + pos[0]: the same sum over the batch
        dpT = ops.transpose_bf16(de16, gather=(S - 1, S, 1))  # patch rows only: [D, B*(S-1) padded]
        dWp = ops.gemm(dpT, ops.transpose_bf16(cols), out_fp32=True)  # [D, Kpad]
        d_patch_w = (dWp if Kpad == K else dWp[:, :K].contiguous()).view(pw_shape)
        return (None, None, d_patch_w, d_cls, d_pos, d_pre_g, d_pre_b, *lgrads)


class _TextTowerFn(torch.autograd.Function):
    """input_ids -> final_layer_norm(hidden)[:, 0] as fp32 [B, D] (HF:531-562; token-0 pooling of model_m.py:102)."""

    @staticmethod
    def forward(ctx, tw, input_ids, key_mask, *params):
        tok, pos, fin_g, fin_b = params[:4]
        lp = list(params[4:])
        keep = any(ctx.needs_input_grad[3:])
        B, S = input_ids.shape
        D = tw.Dt
        x0 = ops.text_embed(input_ids, tok, pos)
        x, saved = _encoder_fwd(x0, lp, tw.Lt, B, S, tw.Ht, tw.eps_t, True, key_mask, keep)
        out = ops.layernorm_rows_f32(x, fin_g, fin_b, tw.eps_t, rows=B, ldx=S * D)
        if keep:
            ctx.tw, ctx.saved, ctx.lp = tw, saved, lp
            ctx.head = (x, fin_g, input_ids, key_mask, tuple(tok.shape), tuple(pos.shape))
        return out

    @staticmethod
    def backward(ctx, d_rows):
        tw = ctx.tw
        x, fin_g, input_ids, key_mask, tok_shape, pos_shape = ctx.head
        B, S = input_ids.shape
        D = tw.Dt
        dx32 = torch.zeros((B * S, D), device=d_rows.device, dtype=f32)
        d_fin_g, d_fin_b = ops.layernorm_bwd(d_rows.contiguous(), x.view(B, S * D)[:, :D], fin_g, tw.eps_t,
                                             dx_f32=dx32.view(B, S * D)[:, :D])
        dx32, lgrads = _encoder_bwd(dx32, ctx.saved, ctx.lp, B, S, tw.Ht, tw.eps_t, True, key_mask, tw.layer_grad_sink)
        ctx.saved = None
        d_tok = torch.zeros(tok_shape, device=dx32.device, dtype=f32)
        ops.embed_scatter_add(dx32, input_ids.reshape(-1), d_tok)
        d_pos = torch.zeros(pos_shape, device=dx32.device, dtype=f32)
        d_pos[:S].copy_(ops.colsum_f32(dx32.view(B, S * D)).view(S, D))
        return (None, None, None, d_tok, d_pos, d_fin_g, d_fin_b, *lgrads)


class _LinearF32TrainFn(torch.autograd.Function):
    """y = x W^T with a TRAINABLE bias-free weight (visual_projection / text_projection, HF:784-785): dx and dW."""

    @staticmethod
    def forward(ctx, x, W):
        R, K = x.shape
        Nn = W.shape[0]
        y = torch.empty((R, Nn), device=x.device, dtype=f32)
        N.check(N.load().vlmclip_linear_f32(N.ptr(x), x.stride(0), N.ptr(W), None, N.ptr(y), R, Nn, K, N.stream()),
                "vlmclip_linear_f32")
        ctx.save_for_backward(x, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W = ctx.saved_tensors
        R, K = x.shape
        Nn = W.shape[0]
        dy = dy.contiguous()
        dx = dW = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((R, K), device=dy.device, dtype=f32)
            N.check(N.load().vlmclip_linear_f32_dgrad(N.ptr(dy), N.ptr(W), N.ptr(dx), R, Nn, K, N.stream()),
                    "vlmclip_linear_f32_dgrad")
        if ctx.needs_input_grad[1]:
            dW = ops.linear_f32_wgrad(dy, x)
        return dx, dW


def linear_f32_trainable(x, W):
    ops._req(x.dtype == f32 and x.dim() == 2 and x.is_contiguous(), "linear_f32_trainable: x must be contiguous fp32 [R, K]")
    ops._req(W.dtype == f32 and W.is_contiguous() and W.shape[1] == x.shape[1], "linear_f32_trainable: W must be fp32 [N, K]")
    return _LinearF32TrainFn.apply(x, W)


def track_m_unused_parameters(clip):
    """CLIP parameters Track M never reads (no gradient; torch.optim skips them, and so must a fused optimiser):
    vision post_layernorm (model_m.py:110-122 takes last_hidden_state, HF:684-686 applies it to the pooler only)."""
    named = dict(clip.named_parameters())
    return [named["vision_model.post_layernorm.weight"], named["vision_model.post_layernorm.bias"]]


class TrainableClipTowers:
    """Both towers executed from the live parameters of `clip`, differentiable w.r.t. every one of them."""

    def __init__(self, clip):
        N.load()
        p = next(clip.parameters())
        if p.device.type != "cuda":
            raise N.NativeError("TrainableClipTowers needs a CUDA model: the towers only run on the sm_100a library")
        self.clip = clip
        # data parallel full fine-tune: trainer.BucketedGradAllReduce installs itself here (see _encoder_bwd)
        self.layer_grad_sink = None
        cfg = clip.config
        vc, tc = cfg.vision_config, cfg.text_config
        self.eps_v, self.eps_t = float(vc.layer_norm_eps), float(tc.layer_norm_eps)
        self.Dv, self.Dt = vc.hidden_size, tc.hidden_size
        self.Hv, self.Ht = vc.num_attention_heads, tc.num_attention_heads
        self.Lv, self.Lt = vc.num_hidden_layers, tc.num_hidden_layers
        self.patch, self.image = vc.patch_size, vc.image_size
        self.Sv = (vc.image_size // vc.patch_size) ** 2 + 1
        self.max_pos = tc.max_position_embeddings
        if self.Dv // self.Hv != 64 or self.Dt // self.Ht != 64:
            raise ValueError("the attention kernels are specialised for head_dim = 64 (all OpenAI CLIP towers)")
        if vc.hidden_act != "quick_gelu" or tc.hidden_act != "quick_gelu":
            raise ValueError("only quick_gelu towers are supported (OpenAI CLIP)")
        named = dict(clip.named_parameters())
        vm, tm = "vision_model.", "text_model."
        self.v_params = [named[vm + "embeddings.patch_embedding.weight"], named[vm + "embeddings.class_embedding"],
                         named[vm + "embeddings.position_embedding.weight"], named[vm + "pre_layrnorm.weight"],
                         named[vm + "pre_layrnorm.bias"]]
        for l in range(self.Lv):
            self.v_params += [named[f"{vm}encoder.layers.{l}.{n}"] for n in _LAYER_PARAMS]
        self.t_params = [named[tm + "embeddings.token_embedding.weight"], named[tm + "embeddings.position_embedding.weight"],
                         named[tm + "final_layer_norm.weight"], named[tm + "final_layer_norm.bias"]]
        for l in range(self.Lt):
            self.t_params += [named[f"{tm}encoder.layers.{l}.{n}"] for n in _LAYER_PARAMS]
        for q in self.v_params + self.t_params:
            if q.dtype != f32 or not q.is_contiguous():
                raise N.NativeError("full fine-tuning keeps the CLIP master weights as contiguous fp32 parameters")
        self.visual_projection = named["visual_projection.weight"]
        self.text_projection = named["text_projection.weight"]

    def unused_parameters(self):
        return track_m_unused_parameters(self.clip)

    def vision_cls(self, pixel_values: torch.Tensor) -> torch.Tensor:
        if pixel_values.dim() != 4 or pixel_values.shape[1] != 3:
            raise ValueError("pixel_values must be [B, 3, H, W]")
        if pixel_values.shape[2] != self.image or pixel_values.shape[3] != self.image:
            raise ValueError(f"Input image size ({pixel_values.shape[2]}*{pixel_values.shape[3]}) doesn't match model "
                             f"({self.image}*{self.image}).")
        if pixel_values.dtype not in (f32, bf16):
            pixel_values = pixel_values.float()
        return _VisionTowerFn.apply(self, pixel_values.contiguous(), *self.v_params)

    def text_tok0(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor]) -> torch.Tensor:
        if input_ids.dim() != 2:
            raise ValueError("input_ids must be [B, S]")
        if input_ids.shape[1] > self.max_pos:
            raise ValueError(f"Sequence length must be less than max_position_embeddings (got `sequence length`: "
                             f"{input_ids.shape[1]} and max_position_embeddings: {self.max_pos}")
        key_mask = None
        if attention_mask is not None:
            key_mask = (attention_mask != 0).to(torch.uint8).contiguous()
        return _TextTowerFn.apply(self, input_ids.to(torch.int64).contiguous(), key_mask, *self.t_params)
