"""Trainable path of SharedMHSAttentionAdapter (reference: adapter/clip_adapter.py:99-128 differentiated by autograd).

One `torch.autograd.Function` for the whole module, fp32 like the rest of the trainable path, hand-written backward:

    h  = text_proj(t)                       e  = image_proj(table)
    hq = norm2(h)                           kv = norm1(e)
    q  = Wq hq + bq                         k = Wk kv + bk,  v = Wv kv + bv        (nn.MultiheadAttention in_proj)
    a  = softmax(q k^T / 8) -> dropout -> . v          (8 heads of 64, attention dropout on the probabilities)
    h2 = hq + out_proj(a)
    y  = h2 + dropout(W2 gelu(W1 norm3(h2) + b1) + b2)

`t` holds ONE row per caption (Track M keeps token 0 only, model_m.py:102, and every text row attends to the table
on its own), `table` is the [S, 768] vision position table shared by the batch (model_m.py:93-96).  Dense products:
vlmclip_linear_f32 / _dgrad / _wgrad; everything else: csrc/shared_adapter.cu.  Dropout masks are drawn with torch's
generator (scaled by 1/(1-p)) and applied inside the kernels; pass p = 0 (or eval mode) for a deterministic run.
"""
from __future__ import annotations

import torch

from .. import _native as N
from .. import ops

f32 = torch.float32


def _lin(x, W, b):
    y = torch.empty((x.shape[0], W.shape[0]), device=x.device, dtype=f32)
    N.check(N.load().vlmclip_linear_f32(N.ptr(x), x.stride(0), N.ptr(W), N.ptr(b), N.ptr(y), x.shape[0], W.shape[0],
                                        W.shape[1], N.stream()), "vlmclip_linear_f32")
    return y


def _dgrad(dy, W):
    dx = torch.empty((dy.shape[0], W.shape[1]), device=dy.device, dtype=f32)
    N.check(N.load().vlmclip_linear_f32_dgrad(N.ptr(dy), N.ptr(W), N.ptr(dx), dy.shape[0], W.shape[0], W.shape[1],
                                              N.stream()), "vlmclip_linear_f32_dgrad")
    return dx


def _lin_bwd(dy, x, W, need_dx=True):
    """(dx, dW, db) of y = x W^T + b."""
    return (_dgrad(dy, W) if need_dx else None), ops.linear_f32_wgrad(dy, x), ops.colsum_f32(dy)


def _ln(x, g, b, eps):
    M, D = x.shape
    y = torch.empty((M, D), device=x.device, dtype=f32)
    st = torch.empty((M, 2), device=x.device, dtype=f32)
    N.check(N.load().vlmclip_layernorm_f32(N.ptr(x), x.stride(0), N.ptr(g), N.ptr(b), N.ptr(y), N.ptr(st), M, D, float(eps),
                                           N.stream()), "vlmclip_layernorm_f32")
    return y, st


def _ln_bwd(dy, x, st, g, dres=None):
    M, D = x.shape
    dx = torch.empty((M, D), device=x.device, dtype=f32)
    dg = torch.empty((D,), device=x.device, dtype=f32)
    db = torch.empty((D,), device=x.device, dtype=f32)
    N.check(N.load().vlmclip_layernorm_f32_bwd(N.ptr(dy), N.ptr(x), x.stride(0), N.ptr(st), N.ptr(g), N.ptr(dres), N.ptr(dx),
                                               N.ptr(dg), N.ptr(db), M, D, N.stream()), "vlmclip_layernorm_f32_bwd")
    return dx, dg, db


def _fma(a, mask=None, b=None):
    y = torch.empty_like(a)
    N.check(N.load().vlmclip_fma_mask_f32(N.ptr(a), N.ptr(mask), N.ptr(b), N.ptr(y), a.numel(), N.stream()),
            "vlmclip_fma_mask_f32")
    return y


def _dropout_mask(shape, p, device):
    if p <= 0.0:
        return None
    return (torch.rand(shape, device=device) >= p).to(f32).mul_(1.0 / (1.0 - p))  # mask generation, not activation math


class SharedAdapterTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, table, heads, p_drop, eps, Wt, bt, Wi, bi, Win, bin_, Wo, bo, g1, b1, g2, b2, g3, b3, W1, bm1, W2, bm2):
        lib = N.load()
        B, D = t.shape[0], Wt.shape[0]
        S = table.shape[0]
        scale = (D // heads) ** -0.5
        Wq, Wk, Wv = Win[:D], Win[D:2 * D], Win[2 * D:]
        bq, bk, bv = bin_[:D], bin_[D:2 * D], bin_[2 * D:]
        h = _lin(t, Wt, bt)
        e = _lin(table, Wi, bi)
        kv, st1 = _ln(e, g1, b1, eps[0])
        hq, st2 = _ln(h, g2, b2, eps[1])
        q, k, v = _lin(hq, Wq, bq), _lin(kv, Wk, bk), _lin(kv, Wv, bv)
        pmask = _dropout_mask((B, heads, S), p_drop, t.device)
        p = torch.empty((B, heads, S), device=t.device, dtype=f32)
        att = torch.empty((B, D), device=t.device, dtype=f32)
        N.check(lib.vlmclip_attn1q_f32_fwd(N.ptr(q), N.ptr(k), N.ptr(v), D, N.ptr(pmask), N.ptr(p), N.ptr(att), B, S, heads,
                                           float(scale), N.stream()), "vlmclip_attn1q_f32_fwd")
        h2 = _fma(_lin(att, Wo, bo), None, hq)
        z, st3 = _ln(h2, g3, b3, eps[2])
        a1 = _lin(z, W1, bm1)
        z1 = torch.empty_like(a1)
        N.check(lib.vlmclip_gelu_f32(N.ptr(a1), N.ptr(z1), a1.numel(), N.stream()), "vlmclip_gelu_f32")
        omask = _dropout_mask((B, D), p_drop, t.device)
        y = _fma(_lin(z1, W2, bm2), omask, h2)
        ctx.save_for_backward(t, table, Wt, Wi, Win, Wo, g1, g2, g3, W1, W2)
        ctx.inter = (h, e, kv, st1, hq, st2, q, k, v, pmask, p, att, h2, z, st3, a1, z1, omask)
        ctx.cfg = (heads, scale)
        return y

    @staticmethod
    def backward(ctx, dy):
        t, table, Wt, Wi, Win, Wo, g1, g2, g3, W1, W2 = ctx.saved_tensors
        h, e, kv, st1, hq, st2, q, k, v, pmask, p, att, h2, z, st3, a1, z1, omask = ctx.inter
        heads, scale = ctx.cfg
        lib = N.load()
        B, D = h.shape
        S = e.shape[0]
        Wq, Wk, Wv = Win[:D], Win[D:2 * D], Win[2 * D:]
        dy = dy.contiguous()
        # ---- y = h2 + mask * (W2 z1 + b2) ----
        dz2 = _fma(dy, omask) if omask is not None else dy
        dz1, dW2, dbm2 = _lin_bwd(dz2, z1, W2)
        da1 = torch.empty_like(a1)
        N.check(lib.vlmclip_gelu_f32_bwd(N.ptr(a1), N.ptr(dz1), N.ptr(da1), a1.numel(), N.stream()), "vlmclip_gelu_f32_bwd")
        dz, dW1, dbm1 = _lin_bwd(da1, z, W1)
        dh2, dg3, db3 = _ln_bwd(dz, h2, st3, g3, dres=dy)  # + the skip connection's gradient
        # ---- h2 = hq + out_proj(att) ----
        datt, dWo, dbo = _lin_bwd(dh2, att, Wo)
        ds = torch.empty((B, heads, S), device=dy.device, dtype=f32)
        dq = torch.empty((B, D), device=dy.device, dtype=f32)
        dk = torch.empty((S, D), device=dy.device, dtype=f32)
        dv = torch.empty((S, D), device=dy.device, dtype=f32)
        N.check(lib.vlmclip_attn1q_f32_bwd(N.ptr(datt), N.ptr(q), N.ptr(k), N.ptr(v), D, N.ptr(p), N.ptr(pmask), N.ptr(ds),
                                           N.ptr(dq), N.ptr(dk), N.ptr(dv), B, S, heads, float(scale), N.stream()),
                "vlmclip_attn1q_f32_bwd")
        dhq_q, dWq, dbq = _lin_bwd(dq, hq, Wq)
        dkv_k, dWk, dbk = _lin_bwd(dk, kv, Wk)
        dkv_v, dWv, dbv = _lin_bwd(dv, kv, Wv)
        dhq = _fma(dhq_q, None, dh2)
        dkv = _fma(dkv_k, None, dkv_v)
        dh, dg2, db2 = _ln_bwd(dhq, h, st2, g2)
        de, dg1, db1 = _ln_bwd(dkv, e, st1, g1)
        dt, dWt, dbt = _lin_bwd(dh, t, Wt, need_dx=ctx.needs_input_grad[0])
        dtab, dWi, dbi = _lin_bwd(de, table, Wi, need_dx=ctx.needs_input_grad[1])
        dWin = torch.cat([dWq, dWk, dWv], 0)  # nn.MultiheadAttention keeps q/k/v in one in_proj parameter (a copy)
        dbin = torch.cat([dbq, dbk, dbv], 0)
        ctx.inter = None
        return (dt, dtab, None, None, None, dWt, dbt, dWi, dbi, dWin, dbin, dWo, dbo, dg1, db1, dg2, db2, dg3, db3, dW1,
                dbm1, dW2, dbm2)


def shared_adapter_train(mod, t, table):
    """mod: SharedMHSAttentionAdapter; t fp32 [B, text_in] (one row per caption); table fp32 [S, image_in]."""
    ops._req(t.dtype == f32 and t.dim() == 2 and t.is_contiguous(), "shared adapter: text rows must be contiguous fp32 [B, D]")
    ops._req(table.dtype == f32 and table.dim() == 2 and table.is_contiguous(), "shared adapter: table must be fp32 [S, D]")
    ca = mod.cross_attn
    D = mod.text_proj.out_features
    if D // ca.num_heads != 64:
        raise ValueError("the attention kernels are specialised for head_dim = 64")
    p = float(ca.dropout) if mod.training else 0.0
    if float(mod.mlp[3].p) != float(ca.dropout):
        raise ValueError("attention and MLP dropout differ: the reference uses one `dropout` for both")
    params = (mod.text_proj.weight, mod.text_proj.bias, mod.image_proj.weight, mod.image_proj.bias, ca.in_proj_weight,
              ca.in_proj_bias, ca.out_proj.weight, ca.out_proj.bias, mod.norm1.weight, mod.norm1.bias, mod.norm2.weight,
              mod.norm2.bias, mod.norm3.weight, mod.norm3.bias, mod.mlp[0].weight, mod.mlp[0].bias, mod.mlp[2].weight,
              mod.mlp[2].bias)
    for q in params:
        ops._req(q.dtype == f32 and q.is_contiguous(), "shared adapter: parameters must be contiguous fp32")
    return SharedAdapterTrainFn.apply(t, table, int(ca.num_heads), p, (mod.norm1.eps, mod.norm2.eps, mod.norm3.eps), *params)
