"""B200-native mirrors of the PE-CLIP style adapters (reference: adapter/peclip.py).

TextualAdapter:  up_proj(GELU(down_proj(x))) + x                  (adapter/peclip.py:13-18)  fused fwd + bwd kernel
ContextAdapter / SharedAdapter:  LayerNorm(MHSA(x, x, x) + x)     (adapter/peclip.py:31-34, 45-48)
    forward on the tensor cores: in_proj GEMM -> flash attention -> out_proj GEMM (+residual) -> LayerNorm.
    These two are inference-only in this round (nothing in the reference wires them into a model or a loss).
State-dict keys match the reference (down_proj / up_proj; mhsa.in_proj_weight ... / layer_norm).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _native as N
from .. import ops


class TextualAdapter(nn.Module):
    def __init__(self, input_dim, hidden_dim):
        super(TextualAdapter, self).__init__()
        self.down_proj = nn.Linear(input_dim, hidden_dim)
        self.up_proj = nn.Linear(hidden_dim, input_dim)
        self.gelu = nn.GELU()

    def _run(self, x2d, rows=None, ldx=None):
        return ops.adapter(x2d, self.down_proj.weight, self.down_proj.bias, self.up_proj.weight, self.up_proj.bias,
                           act=N.ACT_GELU_ERF, post=N.POST_RESIDUAL, rows=rows, ldx=ldx)

    def forward(self, x):
        D = self.down_proj.in_features
        h = x if x.dtype in (torch.float32, torch.bfloat16) else x.float()
        h2 = h.reshape(-1, D)
        if not h2.is_contiguous():
            h2 = h2.contiguous()
        return self._run(h2).view(*x.shape[:-1], D)

    def forward_token0(self, hidden_flat, batch: int, seq: int):
        return self._run(hidden_flat, rows=batch, ldx=seq * self.down_proj.in_features)


class _SelfAttentionAdapter(nn.Module):
    def __init__(self, input_dim, num_heads):
        super().__init__()
        self.mhsa = nn.MultiheadAttention(embed_dim=input_dim, num_heads=num_heads, batch_first=True)
        self.layer_norm = nn.LayerNorm(input_dim)
        self._packed = None

    def _pack(self, dev):
        # _version does not move when ops.FusedAdamW updates the arena in place: the optimiser generation does
        key = (self.mhsa.in_proj_weight._version, self.mhsa.out_proj.weight._version, self.mhsa.in_proj_weight.data_ptr(),
               ops.param_generation(), str(dev))
        if self._packed is None or self._packed[0] != key:
            self._packed = (key, self.mhsa.in_proj_weight.detach().to(dev, torch.bfloat16).contiguous(),
                            self.mhsa.out_proj.weight.detach().to(dev, torch.bfloat16).contiguous())
        return self._packed[1], self._packed[2]

    def forward(self, x):
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise N.NativeError(f"{type(self).__name__}: only the inference path has an sm_100a kernel in this round; "
                                "call under torch.no_grad()")
        B, S, D = x.shape
        H = self.mhsa.num_heads
        if D // H != 64:
            raise ValueError("the attention kernel is specialised for head_dim = 64")
        w_in, w_out = self._pack(x.device)
        xb = x.reshape(B * S, D).to(torch.bfloat16).contiguous()
        qkv = ops.gemm(xb, w_in, bias=self.mhsa.in_proj_bias.detach().float())
        att = ops.attention(qkv, B, S, H)
        z = ops.gemm(att, w_out, bias=self.mhsa.out_proj.bias.detach().float(), residual=xb)
        y = ops.layernorm_rows_f32(z, self.layer_norm.weight.detach(), self.layer_norm.bias.detach(),
                                   self.layer_norm.eps)
        return y.view(B, S, D)


class ContextAdapter(_SelfAttentionAdapter):
    """Processes spatial context in image patches (reference: adapter/peclip.py:21-34)."""


class SharedAdapter(_SelfAttentionAdapter):
    """Reference: adapter/peclip.py:37-48."""
