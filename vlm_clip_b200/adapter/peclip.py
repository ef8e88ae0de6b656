"""B200-native mirrors of the PE-CLIP style adapters (reference: adapter/peclip.py).

TextualAdapter:  up_proj(GELU(down_proj(x))) + x                  (adapter/peclip.py:13-18)  fused fwd + bwd kernel
ContextAdapter / SharedAdapter:  LayerNorm(MHSA(x, x, x) + x)     (adapter/peclip.py:31-34, 45-48)
    on the tensor cores: in_proj GEMM -> flash attention -> out_proj GEMM (+residual) -> LayerNorm, and a hand-written
    backward through the same kernels the full-fine-tune path uses (dgrad / wgrad on the tcgen05 GEMM, tensor-core
    attention backward, LayerNorm backward): gradients of all six parameter tensors and of the input.
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


class _SelfAttnAdapterFn(torch.autograd.Function):
    """y = LayerNorm(out_proj(attention(in_proj(x))) + x) on [B, S, D] (adapter/peclip.py:31-34), bf16 tensor-core
    arithmetic with fp32 accumulation like the towers; backward by hand."""

    @staticmethod
    def forward(ctx, x, w_in, b_in, w_out, b_out, gamma, beta, heads, eps):
        B, S, D = x.shape
        bf16 = torch.bfloat16
        xb = x.reshape(B * S, D)
        xb = xb.contiguous() if xb.dtype == bf16 else ops.cast_bf16(xb.float().contiguous())
        Wi, Wo = ops.cast_bf16(w_in.detach().contiguous()), ops.cast_bf16(w_out.detach().contiguous())
        qkv = ops.gemm(xb, Wi, bias=b_in.detach())
        att = ops.attention(qkv, B, S, heads)
        z = ops.gemm(att, Wo, bias=b_out.detach(), residual=xb)
        y = ops.layernorm_rows_f32(z, gamma.detach(), beta.detach(), eps)
        if any(ctx.needs_input_grad):
            ctx.save_for_backward(xb, qkv, att, z, Wi, Wo, gamma)
            ctx.cfg = (B, S, D, int(heads), float(eps), x.dtype)
        return y.view(B, S, D)

    @staticmethod
    def backward(ctx, dy):
        from ..finetune import _dense_bwd  # y = x W^T + b backward on the forward GEMM kernel (transposed operands)

        xb, qkv, att, z, Wi, Wo, gamma = ctx.saved_tensors
        B, S, D, H, eps, x_dtype = ctx.cfg
        M = B * S
        dy = dy.reshape(M, D)
        dy = dy.contiguous() if dy.dtype in (torch.float32, torch.bfloat16) else dy.float().contiguous()
        dz32 = torch.empty((M, D), device=dy.device, dtype=torch.float32)
        dz16 = torch.empty((M, D), device=dy.device, dtype=torch.bfloat16)
        d_gamma, d_beta = ops.layernorm_bwd(dy, z, gamma.detach(), eps, dx_f32=dz32, dx_bf16=dz16)
        d_att, d_wo, d_bo = _dense_bwd(dz16, att, Wo)
        d_qkv = ops.attention_bwd(qkv, att, d_att, B, S, H)
        d_xb, d_wi, d_bi = _dense_bwd(d_qkv, xb, Wi)
        dx = None
        if ctx.needs_input_grad[0]:
            # the skip connection carries dz in fp32; the attention branch joins it
            N.check(N.load().vlmclip_add_bf16_into_f32(N.ptr(dz32), N.ptr(d_xb), dz32.numel(), N.stream()),
                    "vlmclip_add_bf16_into_f32")
            dx = dz32.view(B, S, D)
            if x_dtype != torch.float32:
                dx = dx.to(x_dtype)
        return dx, d_wi, d_bi, d_wo, d_bo, d_gamma, d_beta, None, None


class _SelfAttentionAdapter(nn.Module):
    def __init__(self, input_dim, num_heads):
        super().__init__()
        self.mhsa = nn.MultiheadAttention(embed_dim=input_dim, num_heads=num_heads, batch_first=True)
        self.layer_norm = nn.LayerNorm(input_dim)

    def forward(self, x):
        """x [B, S, D] (fp32 or bf16, CUDA) -> fp32 [B, S, D].  Differentiable w.r.t. x and the module's parameters."""
        if x.dim() != 3:
            raise ValueError("expected [batch, tokens, features] (nn.MultiheadAttention(batch_first=True), adapter/peclip.py:26)")
        D, H = x.shape[-1], self.mhsa.num_heads
        if D != self.mhsa.embed_dim:
            raise RuntimeError(f"embedding dimension {D} does not match the adapter's {self.mhsa.embed_dim}")
        if D // H != 64:
            raise ValueError("the attention kernel is specialised for head_dim = 64")
        m = self.mhsa
        return _SelfAttnAdapterFn.apply(x, m.in_proj_weight, m.in_proj_bias, m.out_proj.weight, m.out_proj.bias,
                                        self.layer_norm.weight, self.layer_norm.bias, H, self.layer_norm.eps)


class ContextAdapter(_SelfAttentionAdapter):
    """Processes spatial context in image patches (reference: adapter/peclip.py:21-34)."""


class SharedAdapter(_SelfAttentionAdapter):
    """Reference: adapter/peclip.py:37-48."""
