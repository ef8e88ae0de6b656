"""Dimensions of the OpenAI CLIP checkpoints the reference uses (`clip_model_name` of
model_m.py:17, model_t.py:38, model_v.py:60), a random-init `CLIPModel` builder for boxes without network access
(SURVEY.md §8c shim 1: no pretrained weights offline) and the algorithmic FLOP count bench.py reports against."""
from __future__ import annotations

from typing import Dict, Optional

# name -> (vision: width, layers, heads, mlp, tokens), (text: ...), projection dim, patch
CLIP_DIMS = {
    "openai/clip-vit-base-patch32": ((768, 12, 12, 3072, 50), (512, 12, 8, 2048, 77), 512, 32),
    "openai/clip-vit-base-patch16": ((768, 12, 12, 3072, 197), (512, 12, 8, 2048, 77), 512, 16),
    "openai/clip-vit-large-patch14": ((1024, 24, 16, 4096, 257), (768, 12, 12, 3072, 77), 768, 14),
}


def clip_config(name: str, vision_layers: Optional[int] = None, text_layers: Optional[int] = None):
    """transformers.CLIPConfig with the dimensions of checkpoint `name` (nothing is downloaded)."""
    from transformers import CLIPConfig

    (vw, vl, vh, vm, _), (tw, tl, th, tm, ts), proj, patch = CLIP_DIMS[name]
    return CLIPConfig(
        text_config=dict(hidden_size=tw, intermediate_size=tm, num_hidden_layers=text_layers or tl, num_attention_heads=th,
                         max_position_embeddings=ts, vocab_size=49408, projection_dim=proj, eos_token_id=2, bos_token_id=0,
                         pad_token_id=1),
        vision_config=dict(hidden_size=vw, intermediate_size=vm, num_hidden_layers=vision_layers or vl,
                           num_attention_heads=vh, image_size=224, patch_size=patch, projection_dim=proj),
        projection_dim=proj,
    )


def random_init_clip(name: str, seed: int = 0, vision_layers: Optional[int] = None, text_layers: Optional[int] = None):
    """Seeded random-init CLIPModel of checkpoint `name`'s architecture: the weight container `CLIPWithAdapters(clip=...)`
    takes when `from_pretrained` cannot reach the hub."""
    import torch
    from transformers import CLIPModel

    torch.manual_seed(seed)
    m = CLIPModel(clip_config(name, vision_layers, text_layers))
    m.eval()
    return m


def flops_per_pair(name: str) -> Dict[str, float]:
    """Algorithmic forward FLOPs per image / caption / pair (2 x MAC over the dense layers, the attention products, the
    patch embedding and the projection; elementwise work excluded, attention counted dense) - SURVEY.md §8d."""
    (vw, vl, _, vm, vs), (tw, tl, _, tm, ts), proj, patch = CLIP_DIMS[name]

    def tower(w, layers, mlp, seq):
        return layers * (2 * seq * (4 * w * w + 2 * w * mlp) + 4 * seq * seq * w)

    img = tower(vw, vl, vm, vs) + 2 * (vs - 1) * 3 * patch * patch * vw + 2 * vw * proj
    txt = tower(tw, tl, tm, ts) + 2 * tw * proj
    return {"image": float(img), "caption": float(txt), "pair": float(img + txt)}
