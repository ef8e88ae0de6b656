"""B200-native mirrors of the reference's sequence adapters (reference: adapter/clip_adapter.py).

Same class names, constructor signatures, attribute names and state-dict keys (`down_project`, `up_project`,
`layer_norm`; fixture test_checkpoints/test_adapter.pt), so checkpoints are interchangeable and
`trainer.py:40-43`'s `"adapter" in name` filter keeps working.  The arithmetic

    y = LayerNorm(up_project(GELU(down_project(x))) + x)          (adapter/clip_adapter.py:17-23, 144-150)

runs in ONE fused CUDA kernel (vlmclip_adapter_fwd) with a matching backward that produces the adapter's own
gradients (vlmclip_adapter_bwd).  Parameters stay fp32 nn.Parameters owned by torch.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _native as N
from .. import ops


class _SeqBottleneckAdapter(nn.Module):
    def __init__(self, hidden_size, adapter_size):
        super().__init__()
        self.down_project = nn.Linear(hidden_size, adapter_size)
        self.activation = nn.GELU()
        self.up_project = nn.Linear(adapter_size, hidden_size)
        self.layer_norm = nn.LayerNorm(hidden_size)

    def _run(self, x2d, rows=None, ldx=None):
        ln = self.layer_norm
        return ops.adapter(x2d, self.down_project.weight, self.down_project.bias, self.up_project.weight,
                           self.up_project.bias, ln.weight, ln.bias, act=N.ACT_GELU_ERF, post=N.POST_RESIDUAL_LN,
                           eps=ln.eps, rows=rows, ldx=ldx)

    def forward(self, hidden_states):
        """hidden_states: [..., hidden_size] (fp32, or the towers' bf16) on CUDA -> fp32, same shape."""
        D = self.down_project.in_features
        if hidden_states.shape[-1] != D:
            raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied (last dim {hidden_states.shape[-1]} != {D})")
        x = hidden_states
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        x2d = x.reshape(-1, D)
        if not x2d.is_contiguous():
            x2d = x2d.contiguous()
        y = self._run(x2d)
        return y.view(*hidden_states.shape[:-1], D)

    def forward_token0(self, hidden_flat, batch: int, seq: int):
        """Result-identical fast path of `self(h)[:, 0, :]` (model_m.py:102,122): the adapter is position-wise, so
        only row 0 of every sequence is evaluated.  hidden_flat: contiguous [batch*seq, hidden] (bf16 or fp32)."""
        D = self.down_project.in_features
        return self._run(hidden_flat, rows=batch, ldx=seq * D)


class TextAdapter(_SeqBottleneckAdapter):
    """Adapter module for the CLIP text encoder (reference: adapter/clip_adapter.py:4-23)."""


class VisionAdapter(_SeqBottleneckAdapter):
    """Adapter module for the CLIP vision encoder (reference: adapter/clip_adapter.py:131-150)."""


class SharedMHSAttentionAdapter(nn.Module):
    """Cross-modal adapter (reference: adapter/clip_adapter.py:69-128).

    Parameter layout and state-dict keys match the reference (text_proj, image_proj, cross_attn, norm1..3, mlp).
    The reference can only execute this module at batch size 1 (model_m.py:96-100 passes a batch-1 K/V to
    nn.MultiheadAttention) and it hard-codes 512/768 widths; it is outside the accelerated hot
    path.  forward() therefore raises instead of silently running a non-native path.
    """

    def __init__(self, text_input_size=512, image_input_size=768, hidden_size=512, num_heads=8, dropout=0.1):
        super().__init__()
        self.text_proj = nn.Linear(text_input_size, hidden_size)
        self.image_proj = nn.Linear(image_input_size, hidden_size)
        self.cross_attn = nn.MultiheadAttention(embed_dim=hidden_size, num_heads=num_heads, dropout=dropout,
                                                batch_first=True)
        self.norm1 = nn.LayerNorm(hidden_size)
        self.norm2 = nn.LayerNorm(hidden_size)
        self.norm3 = nn.LayerNorm(hidden_size)
        self.mlp = nn.Sequential(nn.Linear(hidden_size, hidden_size * 4), nn.GELU(),
                                 nn.Linear(hidden_size * 4, hidden_size), nn.Dropout(dropout))

    def forward(self, hidden_states, encoder_hidden_states):
        raise N.NativeError(
            "SharedMHSAttentionAdapter has no sm_100a kernel yet (cross-attention with unequal sequence lengths); "
            "construct CLIPWithAdapters(use_shared_adapters=False) — the configuration the reference itself can "
            "train (trainer.py:191-195).")
