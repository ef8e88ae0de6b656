"""Rounding-point emulation of the CUDA tower pipeline on the oracle — TEST INFRASTRUCTURE ONLY.

Replays oracle/clip_oracle.py's towers with the roundings the sm_100a kernels perform (bf16 GEMM operands with fp32
accumulation, LayerNorm folded behind the GEMM, bf16 qkv / attention-output / MLP-hidden tensors, bf16-truncated
softmax numerators) so that the error budget of a storage decision (bf16 residual stream vs a two-term hi+lo stream)
can be measured on the CPU, at full depth, before a kernel is written.  `python -m oracle.emulate_bf16` prints the
table DESIGN.md §4 quotes.  Follows HF modeling_clip.py:363-384 (layer), :261-279 (attention), :347-351 (MLP).
"""
from __future__ import annotations

import sys
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import clip_oracle as O

bf16 = torch.bfloat16


def rb(x: torch.Tensor) -> torch.Tensor:
    """round to bf16 (nearest even) and back"""
    return x.to(bf16).float()


def tb(x: torch.Tensor) -> torch.Tensor:
    """truncate to bf16 (what the attention kernel does to the softmax numerators)"""
    return (x.contiguous().view(torch.int32) & -65536).view(torch.float32)


def _fold_gemm(x_op, mean, rstd, W, b, gamma, beta):
    Wf = rb(W * gamma[None, :])
    c = Wf.sum(1)
    d = W @ beta + b
    acc = x_op @ Wf.t()
    return rstd * (acc - mean * c) + d


def _attention(qkv, heads, causal, key_mask, trunc_p=True):
    B, S, D3 = qkv.shape
    D = D3 // 3
    q, k, v = qkv.split(D, dim=-1)
    hd = D // heads
    qh = q.view(B, S, heads, hd).transpose(1, 2)
    kh = k.view(B, S, heads, hd).transpose(1, 2)
    vh = v.view(B, S, heads, hd).transpose(1, 2)
    s = (qh @ kh.transpose(-1, -2)) * hd ** -0.5
    allow = torch.ones(B, 1, S, S, dtype=torch.bool)
    if causal:
        allow = allow & torch.ones(S, S, dtype=torch.bool).tril()
    if key_mask is not None:
        allow = allow & key_mask.bool()[:, None, None, :]
    s = s.masked_fill(~allow, float("-inf"))
    p = torch.exp(s - s.amax(-1, keepdim=True))
    p = tb(p) if trunc_p else rb(p)
    o = (p @ vh) / p.sum(-1, keepdim=True)
    return o.transpose(1, 2).reshape(B, S, D)


def encoder(x: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str, heads: int, causal: bool,
            key_mask: Optional[torch.Tensor], residual: str, collect=None):
    """residual: 'bf16' (round the stream after every residual add), 'hilo' (stream = bf16 hi + bf16 lo; hi is the GEMM
    operand) or 'fp32'."""
    def store(v):
        if residual == "bf16":
            return rb(v)
        if residual == "hilo":
            hi = rb(v)
            return hi + rb(v - hi)
        return v

    x = store(x)
    n = O._num_layers(sd, prefix)
    for l in range(n):
        p = f"{prefix}encoder.layers.{l}."
        mean = x.mean(-1, keepdim=True)
        rstd = torch.rsqrt((x - mean).pow(2).mean(-1, keepdim=True) + 1e-5)
        Wqkv = torch.cat([sd[p + f"self_attn.{n_}_proj.weight"] for n_ in "qkv"], 0)
        bqkv = torch.cat([sd[p + f"self_attn.{n_}_proj.bias"] for n_ in "qkv"], 0)
        qkv = rb(_fold_gemm(rb(x), mean, rstd, Wqkv, bqkv, sd[p + "layer_norm1.weight"], sd[p + "layer_norm1.bias"]))
        att = rb(_attention(qkv, heads, causal, key_mask))
        x = store(att @ rb(sd[p + "self_attn.out_proj.weight"]).t() + sd[p + "self_attn.out_proj.bias"] + x)
        mean = x.mean(-1, keepdim=True)
        rstd = torch.rsqrt((x - mean).pow(2).mean(-1, keepdim=True) + 1e-5)
        h = _fold_gemm(rb(x), mean, rstd, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"], sd[p + "layer_norm2.weight"],
                       sd[p + "layer_norm2.bias"])
        h = rb(h * (0.5 * torch.tanh(0.851 * h) + 0.5))
        x = store(h @ rb(sd[p + "mlp.fc2.weight"]).t() + sd[p + "mlp.fc2.bias"] + x)
        if collect is not None:
            collect.append(x.clone())
    return x


def vision_tower(sd, pix, heads, residual, collect=None):
    p = "vision_model."
    w = sd[p + "embeddings.patch_embedding.weight"]
    D, _, ps, _ = w.shape
    B = pix.shape[0]
    patches = F.unfold(pix.float(), kernel_size=ps, stride=ps).transpose(1, 2)
    x = rb(rb(patches) @ rb(w.reshape(D, -1)).t())
    cls = sd[p + "embeddings.class_embedding"].expand(B, 1, D)
    x = torch.cat([cls, x], dim=1) + sd[p + "embeddings.position_embedding.weight"][None]
    x = O.layer_norm(x, sd[p + "pre_layrnorm.weight"], sd[p + "pre_layrnorm.bias"])
    return encoder(x, sd, p, heads, False, None, residual, collect)


def text_tower(sd, ids, mask, heads, residual, collect=None):
    p = "text_model."
    S = ids.shape[1]
    x = sd[p + "embeddings.token_embedding.weight"][ids] + sd[p + "embeddings.position_embedding.weight"][:S][None]
    x = encoder(x, sd, p, heads, True, mask, residual, collect)
    return O.layer_norm(x, sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"])


def _rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def main(name="openai/clip-vit-base-patch32", batch=8, layers=None):
    torch.set_grad_enabled(False)
    clip = O.build_hf_clip(name, seed=0, vision_layers=layers, text_layers=layers)
    sd = {k: v.detach() for k, v in clip.state_dict().items()}
    d = O.CLIP_DIMS[name]
    pix, ids, mask = O.synthetic_batch(batch, seed=2)
    ids[:, 0] = torch.arange(batch) * 37 + 5
    torch.manual_seed(1)
    ref_v = O.vision_tower(sd, pix, d.vision.heads)
    ref_t = O.text_tower(sd, ids, mask, d.text.heads)
    fi = ref_v[:, 0] @ sd["visual_projection.weight"].t()
    ft = ref_t[:, 0] @ sd["text_projection.weight"].t()
    for scale in (sd["logit_scale"].exp().item(), 100.0):
        ref = O.contrastive_loss(ft, fi, torch.tensor(scale).log())
        print(f"--- {name} B={batch} logit scale {scale:.2f}: oracle loss {ref['loss'].item():.6f}")
        for residual in ("bf16", "hilo", "fp32"):
            v = vision_tower(sd, pix, d.vision.heads, residual)
            t = text_tower(sd, ids, mask, d.text.heads, residual)
            gi = v[:, 0] @ sd["visual_projection.weight"].t()
            gt = t[:, 0] @ sd["text_projection.weight"].t()
            out = O.contrastive_loss(gt, gi, torch.tensor(scale).log())
            lg, lr = out["logits_per_text"], ref["logits_per_text"]
            print(f"  residual {residual:5s}: hidden v {_rel(v, ref_v):.2e} t {_rel(t, ref_t):.2e} | features img "
                  f"{_rel(out['image_features'], ref['image_features']):.2e} txt "
                  f"{_rel(out['text_features'], ref['text_features']):.2e} | logits relL2 {_rel(lg, lr):.2e} "
                  f"max|d|/max|ref| {((lg - lr).abs().max() / lr.abs().max()).item():.2e} | loss d "
                  f"{abs(out['loss'].item() - ref['loss'].item()):.2e} | argmax same "
                  f"{bool(torch.equal(lg.argmax(1), lr.argmax(1)))}")


if __name__ == "__main__":
    main(*(sys.argv[1:2] or ["openai/clip-vit-base-patch32"]), batch=int(sys.argv[2]) if len(sys.argv) > 2 else 8)
