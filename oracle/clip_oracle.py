"""CPU oracle for the CLIP adapter fine-tuning hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module, and only as the checker (or as the timed CPU baseline), never as a compute path of the product.

It is a plain fp32 PyTorch restatement (functional, state-dict driven) of the arithmetic the reference runs:

  backbone   transformers/models/clip/modeling_clip.py ("HF"; pinned 4.51.3 in the reference's uv.lock:1540,
             5.5.0 installed here; the arithmetic of these paths is identical in both):
             vision embeddings HF:202-218, text embeddings HF:234-258, attention HF:261-279 + 300-336,
             MLP HF:347-351 (quick_gelu), encoder layer HF:363-384, text transformer HF:531-589,
             vision transformer HF:667-691, projections / logit_scale HF:784-786.
  adapters   adapter/clip_adapter.py:17-23,144-150 (LN(up(gelu(down x)) + x)), adapter/peclip.py:13-18,31-34,
             45-48, model_t.py:21-22, model_v.py:26-27.
  models     model_m.py:77-176 (token-0 pooling, symmetric InfoNCE), model_t.py:157-187,213-298,
             model_v.py:240-343 (alpha/beta/gamma blends, class-prompt CE, predict paths).
  trainer    trainer.py:39-48,91-99 (clip_grad_norm_ 1.0 + AdamW).

Parity pinning: tests/test_oracle.py checks this file against (a) HF's own CLIPModel on seeded random
weights, (b) golden vectors produced by importing the UNMODIFIED reference modules from /root/reference
(oracle/make_golden.py, fixtures under tests/golden/), (c) the survey's known answers G1/G2.
The reference itself ships no numeric test for this path ("parity unpinned" by the reference; pinned here
against its executed code).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LN_EPS = 1e-5


# ------------------------------------------------------------------------------------------------ configs
@dataclass(frozen=True)
class TowerDims:
    width: int
    layers: int
    heads: int
    mlp: int
    seq: int  # tokens (vision: patches + 1; text: max positions)


@dataclass(frozen=True)
class ClipDims:
    name: str
    vision: TowerDims
    text: TowerDims
    proj: int
    patch: int
    image: int = 224
    vocab: int = 49408


CLIP_DIMS = {
    "openai/clip-vit-base-patch32": ClipDims("B/32", TowerDims(768, 12, 12, 3072, 50), TowerDims(512, 12, 8, 2048, 77), 512, 32),
    "openai/clip-vit-base-patch16": ClipDims("B/16", TowerDims(768, 12, 12, 3072, 197), TowerDims(512, 12, 8, 2048, 77), 512, 16),
    "openai/clip-vit-large-patch14": ClipDims("L/14", TowerDims(1024, 24, 16, 4096, 257), TowerDims(768, 12, 12, 3072, 77), 768, 14),
}


def hf_config(name: str, vision_layers: Optional[int] = None, text_layers: Optional[int] = None):
    """transformers.CLIPConfig with the dims of the OpenAI checkpoint `name` (no download; random init)."""
    from transformers import CLIPConfig

    d = CLIP_DIMS[name]
    return CLIPConfig(
        text_config=dict(hidden_size=d.text.width, intermediate_size=d.text.mlp, num_hidden_layers=text_layers or d.text.layers,
                         num_attention_heads=d.text.heads, max_position_embeddings=77, vocab_size=d.vocab,
                         projection_dim=d.proj, eos_token_id=2, bos_token_id=0, pad_token_id=1),
        vision_config=dict(hidden_size=d.vision.width, intermediate_size=d.vision.mlp,
                           num_hidden_layers=vision_layers or d.vision.layers, num_attention_heads=d.vision.heads,
                           image_size=d.image, patch_size=d.patch, projection_dim=d.proj),
        projection_dim=d.proj,
    )


def build_hf_clip(name: str, seed: int = 0, vision_layers: Optional[int] = None, text_layers: Optional[int] = None):
    """Seeded random-init CLIPModel (the weight container shared by oracle and product; SURVEY.md §8c shim 1)."""
    from transformers import CLIPModel

    torch.manual_seed(seed)
    m = CLIPModel(hf_config(name, vision_layers, text_layers))
    m.eval()
    return m


# ------------------------------------------------------------------------------------------------ primitives
def quick_gelu(x: Tensor) -> Tensor:  # HF:349 via ACT2FN["quick_gelu"]
    return x * torch.sigmoid(1.702 * x)


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = LN_EPS) -> Tensor:
    mu = x.mean(-1, keepdim=True)
    var = (x - mu).pow(2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def linear(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    y = x @ w.t()
    return y if b is None else y + b


def attention_core(q: Tensor, k: Tensor, v: Tensor, heads: int, causal: bool, key_mask: Optional[Tensor]) -> Tensor:
    """softmax(q k^T / sqrt(d) + mask) v per head (HF:261-279).  q,k,v: [B,S,D]; key_mask: [B,S] 1 = attend."""
    B, S, D = q.shape
    hd = D // heads
    qh = q.view(B, S, heads, hd).transpose(1, 2)
    kh = k.view(B, S, heads, hd).transpose(1, 2)
    vh = v.view(B, S, heads, hd).transpose(1, 2)
    scores = (qh @ kh.transpose(-1, -2)) * (hd ** -0.5)
    allow = torch.ones(B, 1, S, S, dtype=torch.bool, device=q.device)
    if causal:
        allow = allow & torch.ones(S, S, dtype=torch.bool, device=q.device).tril()
    if key_mask is not None:
        allow = allow & key_mask.bool()[:, None, None, :]
    scores = scores.masked_fill(~allow, float("-inf"))
    p = torch.softmax(scores.float(), dim=-1)
    p = torch.nan_to_num(p, nan=0.0)  # rows with no visible key -> zeros (documented product behaviour)
    out = p @ vh
    return out.transpose(1, 2).reshape(B, S, D)


def encoder_layer(x: Tensor, sd: Dict[str, Tensor], pre: str, heads: int, causal: bool, key_mask: Optional[Tensor]) -> Tensor:
    """HF:363-384: x += attn(LN1 x); x += fc2(quick_gelu(fc1(LN2 x)))."""
    h = layer_norm(x, sd[pre + "layer_norm1.weight"], sd[pre + "layer_norm1.bias"])
    q = linear(h, sd[pre + "self_attn.q_proj.weight"], sd[pre + "self_attn.q_proj.bias"])
    k = linear(h, sd[pre + "self_attn.k_proj.weight"], sd[pre + "self_attn.k_proj.bias"])
    v = linear(h, sd[pre + "self_attn.v_proj.weight"], sd[pre + "self_attn.v_proj.bias"])
    a = attention_core(q, k, v, heads, causal, key_mask)
    x = x + linear(a, sd[pre + "self_attn.out_proj.weight"], sd[pre + "self_attn.out_proj.bias"])
    h = layer_norm(x, sd[pre + "layer_norm2.weight"], sd[pre + "layer_norm2.bias"])
    h = quick_gelu(linear(h, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"]))
    return x + linear(h, sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])


def _num_layers(sd: Dict[str, Tensor], prefix: str) -> int:
    n = 0
    while f"{prefix}encoder.layers.{n}.layer_norm1.weight" in sd:
        n += 1
    return n


# ------------------------------------------------------------------------------------------------ towers
def vision_tower(sd: Dict[str, Tensor], pixel_values: Tensor, heads: int) -> Tensor:
    """`vision_model(...).last_hidden_state` (HF:667-691): NOT passed through post_layernorm."""
    p = "vision_model."
    w = sd[p + "embeddings.patch_embedding.weight"]  # [D, 3, ps, ps]
    D, _, ps, _ = w.shape
    B = pixel_values.shape[0]
    patches = F.unfold(pixel_values.float(), kernel_size=ps, stride=ps)  # [B, 3*ps*ps, np]
    x = patches.transpose(1, 2) @ w.reshape(D, -1).t()  # conv with kernel = stride (HF:209)
    cls = sd[p + "embeddings.class_embedding"].expand(B, 1, D)
    x = torch.cat([cls, x], dim=1) + sd[p + "embeddings.position_embedding.weight"][None]
    x = layer_norm(x, sd[p + "pre_layrnorm.weight"], sd[p + "pre_layrnorm.bias"])
    for l in range(_num_layers(sd, p)):
        x = encoder_layer(x, sd, f"{p}encoder.layers.{l}.", heads, False, None)
    return x


def text_tower(sd: Dict[str, Tensor], input_ids: Tensor, attention_mask: Optional[Tensor], heads: int) -> Tensor:
    """`text_model(...).last_hidden_state` (HF:531-589): causal AND key-padding mask, then final_layer_norm."""
    p = "text_model."
    S = input_ids.shape[1]
    x = sd[p + "embeddings.token_embedding.weight"][input_ids] + sd[p + "embeddings.position_embedding.weight"][:S][None]
    for l in range(_num_layers(sd, p)):
        x = encoder_layer(x, sd, f"{p}encoder.layers.{l}.", heads, True, attention_mask)
    return layer_norm(x, sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"])


def hf_pooled_image_features(sd: Dict[str, Tensor], pixel_values: Tensor, heads: int) -> Tensor:
    """CLIPModel.get_image_features (HF:829-863, tensor-returning 4.51.3 semantics): post_layernorm(CLS) -> proj."""
    x = vision_tower(sd, pixel_values, heads)[:, 0]
    x = layer_norm(x, sd["vision_model.post_layernorm.weight"], sd["vision_model.post_layernorm.bias"])
    return x @ sd["visual_projection.weight"].t()


def hf_pooled_text_features(sd: Dict[str, Tensor], input_ids: Tensor, attention_mask: Optional[Tensor], heads: int) -> Tensor:
    """CLIPModel.get_text_features (HF:793-827): state at the EOS position (argmax of ids when eos_token_id == 2)."""
    x = text_tower(sd, input_ids, attention_mask, heads)
    pooled = x[torch.arange(x.shape[0]), input_ids.to(torch.int).argmax(dim=-1)]
    return pooled @ sd["text_projection.weight"].t()


# ------------------------------------------------------------------------------------------------ adapters
def bottleneck(x: Tensor, W1: Tensor, b1: Tensor, W2: Tensor, b2: Tensor, act: str) -> Tensor:
    h = linear(x, W1, b1)
    h = F.gelu(h) if act == "gelu" else torch.relu(h)  # nn.GELU() = exact erf form
    return linear(h, W2, b2)


def seq_adapter(x: Tensor, a: Dict[str, Tensor]) -> Tensor:
    """TextAdapter / VisionAdapter (adapter/clip_adapter.py:17-23,144-150)."""
    u = bottleneck(x, a["down_project.weight"], a["down_project.bias"], a["up_project.weight"], a["up_project.bias"], "gelu")
    return layer_norm(u + x, a["layer_norm.weight"], a["layer_norm.bias"])


def peclip_textual_adapter(x: Tensor, a: Dict[str, Tensor]) -> Tensor:
    """adapter/peclip.py:13-18: up(gelu(down x)) + x (no LayerNorm)."""
    return bottleneck(x, a["down_proj.weight"], a["down_proj.bias"], a["up_proj.weight"], a["up_proj.bias"], "gelu") + x


def blend_adapter(x: Tensor, a: Dict[str, Tensor], ratio: float) -> Tensor:
    """model_t.py:163-169 / model_v.py:280-286: r*fc2(relu(fc1 x)) + (1-r)*x, re-normalised."""
    u = bottleneck(x, a["fc1.weight"], a["fc1.bias"], a["fc2.weight"], a["fc2.bias"], "relu")
    f = ratio * u + (1 - ratio) * x
    return f / f.norm(dim=-1, keepdim=True)


# ------------------------------------------------------------------------------------------------ Track M
def model_m_text_features(sd, heads_t, input_ids, attention_mask, text_adapter=None) -> Tensor:
    """model_m.py:77-105 without shared adapters: tower -> adapter on ALL tokens -> token 0 (BOS!) -> projection."""
    x = text_tower(sd, input_ids, attention_mask, heads_t)
    if text_adapter is not None:
        x = seq_adapter(x, text_adapter)
    return x[:, 0, :] @ sd["text_projection.weight"].t()


def model_m_image_features(sd, heads_v, pixel_values, vision_adapter=None) -> Tensor:
    """model_m.py:107-125: pre-post_layernorm states -> adapter -> CLS -> projection."""
    x = vision_tower(sd, pixel_values, heads_v)
    if vision_adapter is not None:
        x = seq_adapter(x, vision_adapter)
    return x[:, 0, :] @ sd["visual_projection.weight"].t()


def contrastive_loss(text_features: Tensor, image_features: Tensor, logit_scale: Tensor):
    """model_m.py:146-171.  Returns the same 5-key dict as the reference."""
    t = text_features / text_features.norm(dim=-1, keepdim=True)
    i = image_features / image_features.norm(dim=-1, keepdim=True)
    lpt = (t @ i.t()) * logit_scale.exp()
    lpi = lpt.t()
    labels = torch.arange(t.shape[0], device=t.device)
    loss = (F.cross_entropy(lpt, labels) + F.cross_entropy(lpi, labels)) / 2
    return {"loss": loss, "text_features": t, "image_features": i, "logits_per_text": lpt, "logits_per_image": lpi}


def contrastive_loss_local_rows(t_all: Tensor, i_all: Tensor, t_loc: Tensor, i_loc: Tensor, row0: int, logit_scale: Tensor):
    """Data-parallel statement of the same loss: the global loss with this rank's rows being the
    differentiable leaves t_loc / i_loc; its gradient is dL_global/d(local rows)."""
    n = t_loc.shape[0]
    t = torch.cat([t_all[:row0].detach(), t_loc, t_all[row0 + n:].detach()], 0)
    i = torch.cat([i_all[:row0].detach(), i_loc, i_all[row0 + n:].detach()], 0)
    return contrastive_loss(t, i, logit_scale)["loss"]


def mhsa_adapter(x: Tensor, a: Dict[str, Tensor], heads: int) -> Tensor:
    """ContextAdapter / SharedAdapter (adapter/peclip.py:31-34,45-48): LayerNorm(MHSA(x, x, x) + x)."""
    D = x.shape[-1]
    qkv = linear(x, a["mhsa.in_proj_weight"], a["mhsa.in_proj_bias"])
    q, k, v = qkv.split(D, dim=-1)
    att = attention_core(q, k, v, heads, False, None)
    y = linear(att, a["mhsa.out_proj.weight"], a["mhsa.out_proj.bias"])
    return layer_norm(y + x, a["layer_norm.weight"], a["layer_norm.bias"])


def shared_mhs_adapter(t: Tensor, table: Tensor, a: Dict[str, Tensor], heads: int = 8) -> Tensor:
    """SharedMHSAttentionAdapter.forward in eval mode (adapter/clip_adapter.py:99-128): t [B, T, 512] text states,
    table [1 or B, S, 768] image-side states (model_m.py:93-96 passes the vision position table, batch 1)."""
    h = linear(t, a["text_proj.weight"], a["text_proj.bias"])
    e = linear(table, a["image_proj.weight"], a["image_proj.bias"])
    kv = layer_norm(e, a["norm1.weight"], a["norm1.bias"])
    h = layer_norm(h, a["norm2.weight"], a["norm2.bias"])
    D = h.shape[-1]
    wq, wk, wv = a["cross_attn.in_proj_weight"].split(D, dim=0)
    bq, bk, bv = a["cross_attn.in_proj_bias"].split(D, dim=0)
    q, k, v = linear(h, wq, bq), linear(kv, wk, bk), linear(kv, wv, bv)
    B, T, _ = q.shape
    k, v = k.expand(B, -1, -1), v.expand(B, -1, -1)
    hd = D // heads
    qh = q.view(B, T, heads, hd).transpose(1, 2)
    kh = k.reshape(B, -1, heads, hd).transpose(1, 2)
    vh = v.reshape(B, -1, heads, hd).transpose(1, 2)
    p = torch.softmax((qh @ kh.transpose(-1, -2)) * hd ** -0.5, dim=-1)
    att = (p @ vh).transpose(1, 2).reshape(B, T, D)
    h = h + linear(att, a["cross_attn.out_proj.weight"], a["cross_attn.out_proj.bias"])
    z = layer_norm(h, a["norm3.weight"], a["norm3.bias"])
    z = linear(F.gelu(linear(z, a["mlp.0.weight"], a["mlp.0.bias"])), a["mlp.2.weight"], a["mlp.2.bias"])
    return h + z


def model_m_forward(sd, heads_t, heads_v, input_ids, attention_mask, pixel_values, text_adapter, vision_adapter):
    t = model_m_text_features(sd, heads_t, input_ids, attention_mask, text_adapter)
    i = model_m_image_features(sd, heads_v, pixel_values, vision_adapter)
    return contrastive_loss(t, i, sd["logit_scale"])


# ------------------------------------------------------------------------------------------------ Track T / V
def class_prompt_logits(image_features_n: Tensor, class_embeddings_n: Tensor, visual_adapter, text_adapter,
                        alpha: float, beta: float, temperature, context_features=None, context_adapter=None,
                        gamma: float = 0.0) -> Tensor:
    """model_t.py:157-184 / model_v.py:260-343 (eval-mode dropout): temperature * f_img f_txt^T."""
    f_img = blend_adapter(image_features_n, visual_adapter, alpha)
    if context_features is not None and context_adapter is not None:
        f_ctx = blend_adapter(context_features, context_adapter, gamma)
        f_img = (f_img + f_ctx) / 2.0
        f_img = f_img / f_img.norm(dim=-1, keepdim=True)
    f_txt = blend_adapter(class_embeddings_n, text_adapter, beta)
    return temperature * (f_img @ f_txt.t())


def class_prompt_loss(logits: Tensor, labels: Tensor) -> Tensor:
    """nn.CrossEntropyLoss()(logits, labels): int64 class indices or fp32 class probabilities (soft labels)."""
    return F.cross_entropy(logits, labels)


def predict_all_descriptions(image_features_n: Tensor, per_prompt_embeddings_n: Tensor, group: int,
                             visual_adapter, text_adapter, alpha: float, beta: float) -> Tensor:
    """model_t.py:244-298: 100 * f_img f_prompt^T, max over the `group` prompts of a class, softmax."""
    f_img = blend_adapter(image_features_n, visual_adapter, alpha)
    f_txt = blend_adapter(per_prompt_embeddings_n, text_adapter, beta)
    sims = 100.0 * (f_img @ f_txt.t())
    C = per_prompt_embeddings_n.shape[0] // group
    return torch.softmax(sims.view(-1, C, group).max(dim=-1).values, dim=1)


# ------------------------------------------------------------------------------------------------ trainer step
def linear_warmup_lr(step: int, warmup: int, total: int) -> float:
    """Multiplier of transformers.get_linear_schedule_with_warmup (trainer.py:58-62)."""
    if step < warmup:
        return step / max(1, warmup)
    return max(0.0, (total - step) / max(1, total - warmup))


def adamw_clip_reference(params, grads, exp_avg, exp_avg_sq, step: int, lr: float, betas=(0.9, 0.999), eps=1e-8,
                         weight_decay=0.01, max_norm=1.0):
    """clip_grad_norm_(max_norm) then one AdamW step (trainer.py:95-98), written out on flat tensors."""
    norm = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    coef = torch.clamp(max_norm / (norm + 1e-6), max=1.0) if max_norm > 0 else torch.tensor(1.0)
    out = []
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        g = g * coef
        p = p * (1 - lr * weight_decay)
        m = betas[0] * m + (1 - betas[0]) * g
        v = betas[1] * v + (1 - betas[1]) * g * g
        bc1, bc2 = 1 - betas[0] ** step, 1 - betas[1] ** step
        p = p - (lr / bc1) * m / (v.sqrt() / math.sqrt(bc2) + eps)
        out.append((p, m, v))
    return out, norm


# ------------------------------------------------------------------------------------------------ synthetic inputs
def synthetic_batch(batch: int, seed: int = 2, seq: int = 77, image: int = 224, vocab_hi: int = 49406):
    """SURVEY.md §8d inputs: pixels ~ N(0,1); ids ~ U{3..49405} with BOS=49406 first and EOS=49407 last; mask = 1."""
    g = torch.Generator().manual_seed(seed)
    pix = torch.randn(batch, 3, image, image, generator=g)
    ids = torch.randint(3, vocab_hi, (batch, seq), generator=g)
    ids[:, 0] = 49406
    ids[:, -1] = 49407
    mask = torch.ones(batch, seq, dtype=torch.int64)
    return pix, ids, mask


def flops_per_pair(name: str) -> Dict[str, float]:
    """Algorithmic forward FLOPs (2*MAC; GEMMs + attention matmuls + patch embed + projection), SURVEY.md §8d."""
    d = CLIP_DIMS[name]

    def tower(t: TowerDims) -> float:
        per_layer = 2 * t.seq * (4 * t.width * t.width + 2 * t.width * t.mlp) + 4 * t.seq * t.seq * t.width
        return t.layers * per_layer

    img = tower(d.vision) + 2 * (d.vision.seq - 1) * 3 * d.patch * d.patch * d.vision.width + 2 * d.vision.width * d.proj
    txt = tower(d.text) + 2 * d.text.width * d.proj
    return {"image": float(img), "caption": float(txt), "pair": float(img + txt)}
