"""Tensor-level wrappers over the C ABI (include/vlmclip.h) plus the autograd glue of the trainable path.

Everything here runs on CUDA through libvlmclip_b200.so.  torch is used to allocate outputs and workspaces
and to carry autograd edges; no torch operator does arithmetic on the hot path.
"""
from __future__ import annotations

import torch

from . import _native as N

bf16 = torch.bfloat16
f32 = torch.float32

# bench.py sets this to {"gemm": []} for ONE instrumented step: every dense-layer launch is then bracketed by
# CUDA events on the launching stream and recorded as (start, stop, algorithmic FLOPs).
PROFILE = None
# development aid (tools/kernel_bench.py): when set to a list, every wrapper call is bracketed by CUDA events and
# recorded as (name, start, stop, extra)
TRACE = None


def _traced(fn):
    import functools

    @functools.wraps(fn)
    def wrapper(*a, **k):
        tr = TRACE
        if tr is None:
            return fn(*a, **k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **k)
        e1.record()
        shape = tuple(a[0].shape) if a and hasattr(a[0], "shape") else ()
        extra = "x".join(map(str, shape))
        if fn.__name__ == "gemm":
            extra += "->%d%s%s%s" % (a[1].shape[0], " res" if k.get("residual") is not None else "",
                                     " act" if k.get("act") else "", " part" if k.get("stats_part_out") is not None else "")
        tr.append((fn.__name__, e0, e1, extra))
        return out

    return wrapper


def _req(cond: bool, msg: str) -> None:
    if not cond:
        raise ValueError(msg)


# --------------------------------------------------------------------------------------------- dense layers
@_traced
def gemm(a, w, bias=None, residual=None, act=N.ACT_NONE, out=None, out_fp32=False, row_stats=None, col_c=None,
         stats_part_in=None, ln_eps=1e-5, stats_part_out=None, stats_out=None, row_counters=None):
    """out[M,N] = epilogue(a[M,K] @ w[N,K]^T) on tcgen05 tensor cores (vlmclip_gemm_bf16)."""
    _req(a.dtype == bf16 and w.dtype == bf16, "gemm: a and w must be bf16")
    _req(a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1], "gemm: shape mismatch")
    _req(a.stride(1) == 1 and w.stride(1) == 1, "gemm: operands must be row-major (K contiguous)")
    M, K = a.shape
    Nn = w.shape[0]
    if out is None:
        out = torch.empty((M, Nn), device=a.device, dtype=f32 if out_fp32 else bf16)
    _req(out.stride(1) == 1 and out.shape == (M, Nn), "gemm: bad out")
    _req(out.dtype == (f32 if out_fp32 else bf16), "gemm: out dtype")
    if residual is not None:
        _req(residual.dtype == bf16 and residual.shape == (M, Nn) and residual.stride(1) == 1, "gemm: bad residual")
    lib = N.load()
    prof = PROFILE
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    N.check(
        lib.vlmclip_gemm_bf16(
            N.ptr(a), a.stride(0), N.ptr(w), w.stride(0), N.ptr(out), out.stride(0), N.ptr(bias), N.ptr(residual),
            residual.stride(0) if residual is not None else 0, N.ptr(row_stats), N.ptr(col_c), N.ptr(stats_part_in),
            (K // 32) if stats_part_in is not None else 0, float(ln_eps), N.ptr(stats_part_out), N.ptr(stats_out),
            N.ptr(row_counters), M, Nn, K, int(act), 1 if out_fp32 else 0, N.stream()),
        "vlmclip_gemm_bf16")
    if prof is not None:
        e1.record()
        prof["gemm"].append((e0, e1, 2.0 * M * Nn * K))
    return out


@_traced
def gemm_res2(a, w, bias, x2, stats_part_out=None, stats_out=None, row_counters=None, ln_eps=1e-5):
    """Two-term residual update in place (vlmclip_gemm_bf16_res2): x2[0] + x2[1] += a[M,K] @ w[N,K]^T + bias, with
    x2 = bf16 [2, M, N] holding hi = bf16(x) and lo = bf16(x - hi).  Returns x2."""
    _req(a.dtype == bf16 and w.dtype == bf16 and x2.dtype == bf16, "gemm_res2: a, w, x2 must be bf16")
    _req(a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1] and a.stride(1) == 1 and w.stride(1) == 1,
         "gemm_res2: operands must be row-major [M, K] / [N, K]")
    M, K = a.shape
    Nn = w.shape[0]
    _req(x2.dim() == 3 and x2.shape == (2, M, Nn) and x2.stride(2) == 1, "gemm_res2: x2 must be [2, M, N]")
    prof = PROFILE
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    N.check(
        N.load().vlmclip_gemm_bf16_res2(N.ptr(a), a.stride(0), N.ptr(w), w.stride(0), N.ptr(x2), x2.stride(1), x2.stride(0),
                                        N.ptr(bias), N.ptr(stats_part_out), N.ptr(stats_out), N.ptr(row_counters),
                                        float(ln_eps), M, Nn, K, N.stream()),
        "vlmclip_gemm_bf16_res2")
    if prof is not None:
        e1.record()
        prof["gemm"].append((e0, e1, 2.0 * M * Nn * K))
    return x2


_SM_COUNT = {}


def _sm_count(dev) -> int:
    n = _SM_COUNT.get(str(dev))
    if n is None:
        n = _SM_COUNT[str(dev)] = torch.cuda.get_device_properties(dev).multi_processor_count
    return n


@_traced
def gemm_splitk(a, w, planes=None):
    """fp32 out[M,N] = a[M,K] @ w[N,K]^T with the REDUCTION distributed over the SMs (vlmclip_gemm_bf16_splitk + an
    in-order sum of the partial planes): for the weight-gradient shapes, M x N = a few output tiles and K = all tokens."""
    _req(a.dtype == bf16 and w.dtype == bf16 and a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1],
         "gemm_splitk: a [M, K], w [N, K] bf16")
    _req(a.stride(1) == 1 and w.stride(1) == 1, "gemm_splitk: operands must be row-major (K contiguous)")
    M, K = a.shape
    Nn = w.shape[0]
    if planes is None:
        # work items = planes x (256 x 256 pair tiles) on sms / 2 CTA pairs: two per pair when the tiles alone are fewer
        tiles = ((M + 255) // 256) * ((Nn + 255) // 256)
        planes = max(1, min(_sm_count(a.device) // tiles, (K + 63) // 64 // 8))
    if planes <= 1:
        return gemm(a, w, out_fp32=True)
    stride = (M * Nn + 3) // 4 * 4
    parts = torch.empty((planes, stride), device=a.device, dtype=f32)
    lib = N.load()
    prof = PROFILE
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    used = lib.vlmclip_gemm_bf16_splitk(N.ptr(a), a.stride(0), N.ptr(w), w.stride(0), N.ptr(parts), Nn, stride, int(planes), M, Nn,
                                        K, N.stream())
    if used <= 0:
        N.check(used if used < 0 else -1, "vlmclip_gemm_bf16_splitk")
    if prof is not None:
        e1.record()
        prof["gemm"].append((e0, e1, 2.0 * M * Nn * K))
    out = torch.empty((M, Nn), device=a.device, dtype=f32)
    _req((M * Nn) % 4 == 0, "gemm_splitk: M * N must be a multiple of 4")
    N.check(lib.vlmclip_sum_planes_f32(N.ptr(parts), stride, int(used), N.ptr(out), M * Nn, N.stream()), "vlmclip_sum_planes_f32")
    return out


@_traced
def gemm_atb_splitk(at, bt, planes=None):
    """fp32 out[M,N] = at[K,M]^T @ bt[K,N] (both bf16, row-major, K = rows) on the tcgen05 GEMM with MN-major operand
    descriptors and a split reduction: the weight gradient dW = dY^T X without transposed copies of dY and X."""
    _req(at.dtype == bf16 and bt.dtype == bf16 and at.dim() == 2 and bt.dim() == 2 and at.shape[0] == bt.shape[0],
         "gemm_atb_splitk: at [K, M], bt [K, N] bf16")
    _req(at.stride(1) == 1 and bt.stride(1) == 1, "gemm_atb_splitk: operands must be row-major")
    K, M = at.shape
    Nn = bt.shape[1]
    _req(M % 8 == 0 and Nn % 8 == 0, "gemm_atb_splitk: M and N must be multiples of 8")
    if planes is None:
        tiles = ((M + 255) // 256) * ((Nn + 255) // 256)
        planes = max(1, min(_sm_count(at.device) // tiles, (K + 63) // 64 // 8))
    stride = (M * Nn + 3) // 4 * 4
    parts = torch.empty((planes, stride), device=at.device, dtype=f32)
    lib = N.load()
    prof = PROFILE
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    used = lib.vlmclip_gemm_bf16_atb_splitk(N.ptr(at), at.stride(0), N.ptr(bt), bt.stride(0), N.ptr(parts), Nn, stride,
                                            int(planes), M, Nn, K, N.stream())
    if used <= 0:
        N.check(used if used < 0 else -1, "vlmclip_gemm_bf16_atb_splitk")
    if prof is not None:
        e1.record()
        prof["gemm"].append((e0, e1, 2.0 * M * Nn * K))
    if used == 1:
        return parts[0, :M * Nn].view(M, Nn)
    out = torch.empty((M, Nn), device=at.device, dtype=f32)
    N.check(lib.vlmclip_sum_planes_f32(N.ptr(parts), stride, int(used), N.ptr(out), M * Nn, N.stream()), "vlmclip_sum_planes_f32")
    return out


@_traced
def colsum_bf16(x):
    """out[c] = sum_r x[r, c] for bf16 x [R, C] (fp32 result): the bias gradient of a dense layer from dY as it lies."""
    _req(x.dtype == bf16 and x.dim() == 2 and x.stride(1) == 1 and x.shape[1] % 8 == 0, "colsum_bf16: x must be bf16 [R, C], C % 8 == 0")
    R, Cc = x.shape
    lib = N.load()
    out = torch.empty((Cc,), device=x.device, dtype=f32)
    ws = torch.empty((lib.vlmclip_colsum_bf16_slices(R) * Cc,), device=x.device, dtype=f32)
    N.check(lib.vlmclip_colsum_bf16(N.ptr(x), x.stride(0), N.ptr(out), N.ptr(ws), R, Cc, N.stream()), "vlmclip_colsum_bf16")
    return out


@_traced
def layernorm(x, gamma, beta, eps=1e-5, out=None, stats=None):
    _req(x.dtype == bf16 and x.dim() == 2 and x.stride(1) == 1, "layernorm: x must be bf16 [M, D]")
    M, D = x.shape
    if out is None:
        out = torch.empty((M, D), device=x.device, dtype=bf16)
    N.check(
        N.load().vlmclip_layernorm_bf16(N.ptr(x), x.stride(0), N.ptr(out), out.stride(0), N.ptr(gamma), N.ptr(beta),
                                        N.ptr(stats), M, D, float(eps), N.stream()), "vlmclip_layernorm_bf16")
    return out


@_traced
def layernorm_rows_f32(x, gamma, beta, eps=1e-5, rows=None, ldx=None):
    """fp32 LayerNorm of `rows` rows of a bf16 buffer read with row stride `ldx` (e.g. token 0 of every sequence)."""
    _req(x.dtype == bf16 and x.is_contiguous(), "layernorm_rows_f32: x must be contiguous bf16")
    D = gamma.shape[0]
    if rows is None:
        rows, ldx = x.shape[0], x.stride(0)
    out = torch.empty((rows, D), device=x.device, dtype=f32)
    N.check(
        N.load().vlmclip_layernorm_bf16_f32out(N.ptr(x), int(ldx), N.ptr(out), D, N.ptr(gamma), N.ptr(beta), int(rows),
                                               D, float(eps), N.stream()), "vlmclip_layernorm_bf16_f32out")
    return out


@_traced
def ln_partials_to_stats(part, eps=1e-5, out=None):
    """(mean, M2) per 32-column block [M, npart, 2] -> (mean, rstd) per row [M, 2]."""
    M, npart, _ = part.shape
    if out is None:
        out = torch.empty((M, 2), device=part.device, dtype=f32)
    N.check(N.load().vlmclip_ln_partials_to_stats(N.ptr(part), N.ptr(out), M, npart, float(eps), N.stream()),
            "vlmclip_ln_partials_to_stats")
    return out


@_traced
def row_stats(x, eps=1e-5, out=None):
    _req(x.dtype == bf16 and x.dim() == 2 and x.stride(1) == 1, "row_stats: x must be bf16 [M, D]")
    M, D = x.shape
    if out is None:
        out = torch.empty((M, 2), device=x.device, dtype=f32)
    N.check(N.load().vlmclip_row_stats_bf16(N.ptr(x), x.stride(0), N.ptr(out), M, D, float(eps), N.stream()),
            "vlmclip_row_stats_bf16")
    return out


@_traced
def im2col(pixels, patch: int, out=None):
    _req(pixels.dim() == 4 and pixels.shape[1] == 3 and pixels.is_contiguous(), "im2col: pixels must be [B,3,H,W]")
    _req(pixels.dtype in (f32, bf16), "im2col: pixels must be fp32 or bf16")
    B, _, H, W = pixels.shape
    K = 3 * patch * patch
    Kpad = (K + 63) // 64 * 64
    rows = B * (H // patch) * (W // patch)
    if out is None:
        out = torch.empty((rows, Kpad), device=pixels.device, dtype=bf16)
    N.check(
        N.load().vlmclip_im2col_patches(N.ptr(pixels), 1 if pixels.dtype == bf16 else 0, N.ptr(out), B, H, W, patch,
                                        N.stream()), "vlmclip_im2col_patches")
    return out


# ------------------------------------------------------------------------------------------ preprocessing
_RESIZE_TABLES = {}


def _linear_coeffs(src: int, dst: int):
    """Source index and the two 11-bit weights of every destination index: OpenCV's INTER_LINEAR coefficient rule
    (resize.cpp: fx = (float)((dx + 0.5) * scale - 0.5); clamp; cvRound(w * 2048)), float32 arithmetic as there."""
    import numpy as np

    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * (src / dst) - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    lo, hi = s < 0, s >= src - 1
    s = np.where(lo, 0, np.where(hi, src - 1, s))
    f = np.where(lo | hi, np.float32(0), f).astype(np.float32)
    w0 = np.rint((np.float32(1) - f) * np.float32(2048)).astype(np.int32)
    w1 = np.rint(f * np.float32(2048)).astype(np.int32)
    return np.stack([s.astype(np.int32), w0, w1], 1)


def resize_tables(Hs: int, Ws: int, H: int, W: int, device):
    """Device copies of the row / column coefficient tables for an Hs x Ws -> H x W bilinear resize (cached)."""
    key = (Hs, Ws, H, W, str(device))
    t = _RESIZE_TABLES.get(key)
    if t is None:
        t = (torch.from_numpy(_linear_coeffs(Hs, H)).to(device), torch.from_numpy(_linear_coeffs(Ws, W)).to(device))
        _RESIZE_TABLES[key] = t
    return t


@_traced
def preprocess_patches(frames_u8, H: int, W: int, patch: int, mean, std, bgr: bool = False, out=None):
    """uint8 [..., Hs, Ws, 3] decoded frames -> bf16 im2col rows [n*(H/p)*(W/p), Kpad] (resize, /255, normalise fused)."""
    _req(frames_u8.dtype == torch.uint8 and frames_u8.dim() >= 3 and frames_u8.shape[-1] == 3 and frames_u8.is_contiguous(),
         "preprocess_patches: frames must be contiguous uint8 [..., Hs, Ws, 3]")
    Hs, Ws = int(frames_u8.shape[-3]), int(frames_u8.shape[-2])
    n = frames_u8.numel() // (Hs * Ws * 3)
    K = 3 * patch * patch
    Kpad = (K + 63) // 64 * 64
    rows = n * (H // patch) * (W // patch)
    if out is None:
        out = torch.empty((rows, Kpad), device=frames_u8.device, dtype=bf16)
    ytab = xtab = None
    if (Hs, Ws) != (H, W):
        ytab, xtab = resize_tables(Hs, Ws, H, W, frames_u8.device)
    N.check(
        N.load().vlmclip_preprocess_patches(N.ptr(frames_u8), Hs * Ws * 3, Hs, Ws, 1 if bgr else 0, N.ptr(ytab), N.ptr(xtab),
                                            float(mean[0]), float(mean[1]), float(mean[2]), float(std[0]), float(std[1]),
                                            float(std[2]), N.ptr(out), n, H, W, patch, N.stream()),
        "vlmclip_preprocess_patches")
    return out


class _MeanPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, T):
        _req(x.dtype == f32 and x.dim() == 2 and x.is_contiguous() and x.shape[0] % T == 0, "mean_pool: x must be fp32 [B*T, P]")
        B, P = x.shape[0] // T, x.shape[1]
        y = torch.empty((B, P), device=x.device, dtype=f32)
        N.check(N.load().vlmclip_mean_pool(N.ptr(x), N.ptr(y), B, T, P, N.stream()), "vlmclip_mean_pool")
        ctx.dims = (B, T, P)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, T, P = ctx.dims
        dy = dy.contiguous()
        dx = torch.empty((B * T, P), device=dy.device, dtype=f32)
        N.check(N.load().vlmclip_mean_pool_bwd(N.ptr(dy), N.ptr(dx), B, T, P, N.stream()), "vlmclip_mean_pool_bwd")
        return dx, None


def mean_pool(x, T: int):
    """[B*T, P] -> [B, P]: mean over the T consecutive rows of each clip."""
    return _MeanPool.apply(x, int(T))


@_traced
def vision_embed_ln(patch, cls, pos, gamma, beta, B: int, S: int, eps=1e-5, out=None, out_lo=None):
    """`out_lo` (optional, bf16 [B*S, D]): second term of the two-term residual stream, bf16(value - out)."""
    D = pos.shape[1]
    _req(patch.dtype in (f32, bf16) and patch.shape == (B * (S - 1), D) and patch.is_contiguous(),
         "vision_embed_ln: patch must be fp32 or bf16 [B*(S-1), D]")
    if out is None:
        out = torch.empty((B * S, D), device=pos.device, dtype=bf16)
    N.check(
        N.load().vlmclip_vision_embed_ln(N.ptr(patch), 1 if patch.dtype == bf16 else 0, N.ptr(cls), N.ptr(pos), N.ptr(gamma), N.ptr(beta),
                                         N.ptr(out), N.ptr(out_lo), B, S, D, float(eps), N.stream()), "vlmclip_vision_embed_ln")
    return out


@_traced
def text_embed(ids, tok, pos, out=None, out_lo=None):
    _req(ids.dtype == torch.int64 and ids.dim() == 2 and ids.is_contiguous(), "text_embed: ids must be int64 [B,S]")
    B, S = ids.shape
    V, D = tok.shape
    _req(pos.shape[0] >= S, "text_embed: sequence longer than the position table")
    if out is None:
        out = torch.empty((B * S, D), device=ids.device, dtype=bf16)
    N.check(
        N.load().vlmclip_text_embed(N.ptr(ids), N.ptr(tok), 1 if tok.dtype == bf16 else 0, N.ptr(pos), N.ptr(out),
                                    N.ptr(out_lo), B, S, D, V, N.stream()), "vlmclip_text_embed")
    return out


@_traced
def attention(qkv, B: int, S: int, H: int, causal=False, key_mask=None, scale=None, out=None):
    _req(qkv.dtype == bf16 and qkv.shape == (B * S, 3 * H * 64) and qkv.is_contiguous(),
         "attention: qkv must be contiguous bf16 [B*S, 3*H*64]")
    if key_mask is not None:
        _req(key_mask.dtype == torch.uint8 and key_mask.shape == (B, S) and key_mask.is_contiguous(),
             "attention: key_mask must be uint8 [B,S]")
    if out is None:
        out = torch.empty((B * S, H * 64), device=qkv.device, dtype=bf16)
    lib = N.load()
    n_ws = 0 if (causal or key_mask is not None) else lib.vlmclip_attention_fwd_workspace(B, S, H)
    ws = torch.empty(n_ws, device=qkv.device, dtype=f32) if n_ws > 0 else None
    N.check(
        lib.vlmclip_attention_fwd_ws(N.ptr(qkv), N.ptr(out), N.ptr(key_mask), N.ptr(ws), B, S, H, 1 if causal else 0,
                                     float(scale if scale is not None else 64 ** -0.5), N.stream()),
        "vlmclip_attention_fwd_ws")
    return out


@_traced
def attention_1q(q, k, v, S: int, H: int, kv_row_stride: int, kv_batch_stride: int = 0, q_stride=None, scale=None):
    """One query row per batch element against S keys (head_dim 64): q bf16 [B, H*64] (row stride q_stride), k / v bf16
    views whose key j of batch b sits at b*kv_batch_stride + j*kv_row_stride (0 = keys shared by the batch)."""
    _req(q.dtype == bf16 and k.dtype == bf16 and v.dtype == bf16, "attention_1q: bf16 operands")
    B = q.shape[0]
    out = torch.empty((B, H * 64), device=q.device, dtype=bf16)
    N.check(
        N.load().vlmclip_attention_1q(N.ptr(q), int(q_stride if q_stride is not None else q.stride(0)), N.ptr(k), N.ptr(v),
                                      int(kv_row_stride), int(kv_batch_stride), N.ptr(out), B, int(S), int(H),
                                      float(scale if scale is not None else 64 ** -0.5), N.stream()),
        "vlmclip_attention_1q")
    return out


@_traced
def gather_rows_f32(x, rows: int, ld: int, D: int, lo=None):
    """y[r] = float(x.flat[r*ld : r*ld + D]) (+ float(lo.flat[...]) for a two-term stream) — the token-0 slice of a
    [B*S, D] bf16 activation."""
    out = torch.empty((rows, D), device=x.device, dtype=f32)
    N.check(N.load().vlmclip_gather_rows2_bf16_to_f32(N.ptr(x), N.ptr(lo), ld, N.ptr(out), rows, D, N.stream()),
            "vlmclip_gather_rows2_bf16_to_f32")
    return out


@_traced
def layernorm_f32(x, gamma, beta, eps=1e-5):
    """fp32 LayerNorm of fp32 rows [R, D] (forward only: the frozen towers' final / post LayerNorm on pooled rows)."""
    _req(x.dtype == f32 and x.dim() == 2 and x.stride(1) == 1, "layernorm_f32: x must be fp32 [R, D]")
    R, D = x.shape
    out = torch.empty((R, D), device=x.device, dtype=f32)
    stats = torch.empty((R, 2), device=x.device, dtype=f32)
    N.check(N.load().vlmclip_layernorm_f32(N.ptr(x), x.stride(0), N.ptr(gamma), N.ptr(beta), N.ptr(out), N.ptr(stats), R, D,
                                           float(eps), N.stream()), "vlmclip_layernorm_f32")
    return out


# ------------------------------------------------------------------------- backbone backward (full fine-tune)
@_traced
def transpose_bf16(src, Rpad=None, gather=None, out=None):
    """out[c, r] = bf16(src[row(r), c]); src [R, C] bf16 or fp32 (row-major, any row stride); out bf16 [C, Rpad] with
    zero-filled columns [R, Rpad) (Rpad defaults to R rounded up to 8: the K of the GEMM that consumes it).
    `gather=(group_dst, group_src, group_off)` reads row (r // group_dst) * group_src + group_off + r % group_dst."""
    _req(src.dim() == 2 and src.stride(1) == 1 and src.dtype in (bf16, f32), "transpose_bf16: src must be bf16/fp32 [R, C]")
    C = src.shape[1]
    gd, gs, go = gather if gather is not None else (0, 0, 0)
    R = src.shape[0] if gather is None else (src.shape[0] // gs) * gd
    if Rpad is None:
        Rpad = (R + 7) // 8 * 8
    if out is None:
        out = torch.empty((C, Rpad), device=src.device, dtype=bf16)
    _req(out.dtype == bf16 and out.shape[0] == C and out.shape[1] >= Rpad and out.stride(1) == 1, "transpose_bf16: bad out")
    N.check(
        N.load().vlmclip_transpose_to_bf16(N.ptr(src), 1 if src.dtype == f32 else 0, src.stride(0), N.ptr(out), out.stride(0),
                                           R, Rpad, C, gd, gs, go, N.stream()), "vlmclip_transpose_to_bf16")
    return out


@_traced
def cast_bf16(src, out=None):
    _req(src.dtype == f32 and src.is_contiguous(), "cast_bf16: src must be contiguous fp32")
    if out is None:
        out = torch.empty(src.shape, device=src.device, dtype=bf16)
    _req(out.dtype == bf16 and out.is_contiguous() and out.numel() == src.numel(), "cast_bf16: bad out")
    N.check(N.load().vlmclip_cast_f32_to_bf16(N.ptr(src), N.ptr(out), src.numel(), N.stream()), "vlmclip_cast_f32_to_bf16")
    return out


@_traced
def rowsum_bf16(x):
    _req(x.dtype == bf16 and x.dim() == 2 and x.stride(1) == 1, "rowsum_bf16: x must be bf16 [R, C]")
    out = torch.empty((x.shape[0],), device=x.device, dtype=f32)
    N.check(N.load().vlmclip_rowsum_bf16(N.ptr(x), x.stride(0), N.ptr(out), x.shape[0], x.shape[1], N.stream()),
            "vlmclip_rowsum_bf16")
    return out


@_traced
def colsum_f32(x):
    _req(x.dtype == f32 and x.dim() == 2 and x.stride(1) == 1, "colsum_f32: x must be fp32 [R, C]")
    out = torch.empty((x.shape[1],), device=x.device, dtype=f32)
    N.check(N.load().vlmclip_colsum_f32(N.ptr(x), x.stride(0), N.ptr(out), x.shape[0], x.shape[1], N.stream()),
            "vlmclip_colsum_f32")
    return out


@_traced
def quick_gelu(a, out=None):
    _req(a.dtype == bf16 and a.is_contiguous(), "quick_gelu: a must be contiguous bf16")
    if out is None:
        out = torch.empty_like(a)
    N.check(N.load().vlmclip_quick_gelu_bf16(N.ptr(a), N.ptr(out), a.numel(), N.stream()), "vlmclip_quick_gelu_bf16")
    return out


@_traced
def quick_gelu_bwd(a, dy, out=None):
    _req(a.dtype == bf16 and dy.dtype == bf16 and a.is_contiguous() and dy.is_contiguous() and a.shape == dy.shape,
         "quick_gelu_bwd: a, dy must be contiguous bf16 of one shape")
    if out is None:
        out = torch.empty_like(a)
    N.check(N.load().vlmclip_quick_gelu_bwd_bf16(N.ptr(a), N.ptr(dy), N.ptr(out), a.numel(), N.stream()),
            "vlmclip_quick_gelu_bwd_bf16")
    return out


@_traced
def layernorm_bwd(dy, x, gamma, eps=1e-5, dres=None, dx_f32=None, dx_bf16=None, param_grads=True):
    """LayerNorm backward.  dy [M, D] bf16 or fp32, x = the LayerNorm's bf16 input [M, D] (both may be row-strided views);
    dres / dx_f32 (fp32) and dx_bf16 share one row stride.  Returns (dgamma, dbeta) (or (None, None))."""
    _req(x.dtype == bf16 and x.dim() == 2 and x.stride(1) == 1, "layernorm_bwd: x must be bf16 [M, D]")
    _req(dy.dtype in (bf16, f32) and dy.shape == x.shape and dy.stride(1) == 1, "layernorm_bwd: dy must match x")
    M, D = x.shape
    lddx = 0
    for t, dt in ((dres, f32), (dx_f32, f32), (dx_bf16, bf16)):
        if t is not None:
            _req(t.dtype == dt and t.shape == x.shape and t.stride(1) == 1, "layernorm_bwd: bad dres / dx")
            _req(lddx in (0, t.stride(0)), "layernorm_bwd: dres, dx_f32 and dx_bf16 must share one row stride")
            lddx = t.stride(0)
    lib = N.load()
    dgamma = dbeta = ws = None
    if param_grads:
        dgamma = torch.empty((D,), device=x.device, dtype=f32)
        dbeta = torch.empty((D,), device=x.device, dtype=f32)
        ws = torch.empty(lib.vlmclip_layernorm_bwd_workspace(M, D), device=x.device, dtype=f32)
    N.check(
        lib.vlmclip_layernorm_bwd(N.ptr(dy), 1 if dy.dtype == f32 else 0, dy.stride(0), N.ptr(x), x.stride(0), N.ptr(gamma),
                                  N.ptr(dres), N.ptr(dx_f32), N.ptr(dx_bf16), lddx, N.ptr(dgamma), N.ptr(dbeta), N.ptr(ws),
                                  M, D, float(eps), N.stream()), "vlmclip_layernorm_bwd")
    return dgamma, dbeta


@_traced
def vision_embed(patch, cls, pos, B: int, S: int, out=None):
    """Vision tokens without pre_layrnorm: bf16 [B*S, D] (the saved LayerNorm input of the training path)."""
    D = pos.shape[1]
    _req(patch.dtype == bf16 and patch.shape == (B * (S - 1), D) and patch.is_contiguous(),
         "vision_embed: patch must be contiguous bf16 [B*(S-1), D]")
    if out is None:
        out = torch.empty((B * S, D), device=pos.device, dtype=bf16)
    N.check(N.load().vlmclip_vision_embed(N.ptr(patch), N.ptr(cls), N.ptr(pos), N.ptr(out), B, S, D, N.stream()),
            "vlmclip_vision_embed")
    return out


@_traced
def embed_scatter_add(d, ids, dtok):
    """dtok[ids[r]] += d[r]: d fp32 [rows, D], ids int64 [rows], dtok fp32 [V, D] (zeroed by the caller)."""
    _req(d.dtype == f32 and d.dim() == 2 and d.stride(1) == 1 and dtok.dtype == f32 and dtok.is_contiguous(),
         "embed_scatter_add: fp32 operands")
    _req(ids.dtype == torch.int64 and ids.is_contiguous() and ids.numel() == d.shape[0], "embed_scatter_add: ids")
    N.check(
        N.load().vlmclip_embed_scatter_add(N.ptr(d), d.stride(0), N.ptr(ids), N.ptr(dtok), d.shape[0], d.shape[1],
                                           dtok.shape[0], N.stream()), "vlmclip_embed_scatter_add")
    return dtok


@_traced
def attention_bwd(qkv, out, dout, B: int, S: int, H: int, causal=False, key_mask=None, scale=None, dqkv=None,
                  simt: bool = False):
    """dqkv of ops.attention.  `simt=True` selects the fp32 SIMT reference kernel instead of the tensor-core pair."""
    for t, w in ((qkv, 3 * H * 64), (out, H * 64), (dout, H * 64)):
        _req(t.dtype == bf16 and t.shape == (B * S, w) and t.is_contiguous(), "attention_bwd: contiguous bf16 operands")
    if key_mask is not None:
        _req(key_mask.dtype == torch.uint8 and key_mask.shape == (B, S) and key_mask.is_contiguous(),
             "attention_bwd: key_mask must be uint8 [B,S]")
    if dqkv is None:
        dqkv = torch.empty_like(qkv)
    lib = N.load()
    ws = None if simt else torch.empty(lib.vlmclip_attention_bwd_workspace(B, S, H), device=qkv.device, dtype=f32)
    N.check(
        lib.vlmclip_attention_bwd(N.ptr(qkv), N.ptr(out), N.ptr(dout), N.ptr(dqkv), N.ptr(key_mask), N.ptr(ws), B, S, H,
                                  1 if causal else 0, float(scale if scale is not None else 64 ** -0.5), N.stream()),
        "vlmclip_attention_bwd")
    return dqkv


@_traced
def linear_f32_wgrad(dy, x):
    """dW[N, K] = dy[R, N]^T x[R, K] (fp32)."""
    _req(dy.dtype == f32 and dy.is_contiguous() and x.dtype == f32 and x.stride(1) == 1 and dy.shape[0] == x.shape[0],
         "linear_f32_wgrad: fp32 dy [R, N], x [R, K]")
    R, Nn = dy.shape
    K = x.shape[1]
    dW = torch.empty((Nn, K), device=dy.device, dtype=f32)
    N.check(N.load().vlmclip_linear_f32_wgrad(N.ptr(dy), N.ptr(x), x.stride(0), N.ptr(dW), R, Nn, K, N.stream()),
            "vlmclip_linear_f32_wgrad")
    return dW


# --------------------------------------------------------------------------------------------- adapters
class _AdapterFn(torch.autograd.Function):
    """Fused bottleneck adapter (vlmclip_adapter_fwd / vlmclip_adapter_bwd)."""

    @staticmethod
    def forward(ctx, x, W1, b1, W2, b2, gamma, beta, hmask, act, post, alpha, eps, ldx, rows):
        D = W1.shape[1]
        A = W1.shape[0]
        x_bf16 = x.dtype == bf16
        y = torch.empty((rows, D), device=W1.device, dtype=f32)
        N.check(
            N.load().vlmclip_adapter_fwd(N.ptr(x), 1 if x_bf16 else 0, ldx, N.ptr(W1), N.ptr(b1), N.ptr(W2),
                                         N.ptr(b2), N.ptr(gamma), N.ptr(beta), N.ptr(hmask), N.ptr(y), rows, D, A,
                                         act, post, float(alpha), float(eps), N.stream()), "vlmclip_adapter_fwd")
        ctx.save_for_backward(x, W1, b1, W2, b2, gamma, beta, hmask)
        ctx.cfg = (act, post, float(alpha), float(eps), ldx, rows, D, A, x_bf16)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W1, b1, W2, b2, gamma, beta, hmask = ctx.saved_tensors
        act, post, alpha, eps, ldx, rows, D, A, x_bf16 = ctx.cfg
        dy = dy.contiguous()
        lib = N.load()
        dev = W1.device
        ws = torch.empty(lib.vlmclip_adapter_bwd_workspace(rows, D, A), device=dev, dtype=f32)
        dW1, db1 = torch.empty_like(W1), torch.empty_like(b1)
        dW2, db2 = torch.empty_like(W2), torch.empty_like(b2)
        has_ln = post == N.POST_RESIDUAL_LN
        dgamma = torch.empty_like(gamma) if has_ln else None
        dbeta = torch.empty_like(beta) if has_ln else None
        need_dx = ctx.needs_input_grad[0]
        dx = torch.empty((rows, D), device=dev, dtype=f32) if need_dx else None
        N.check(
            lib.vlmclip_adapter_bwd(N.ptr(x), 1 if x_bf16 else 0, ldx, N.ptr(W1), N.ptr(b1), N.ptr(W2), N.ptr(b2),
                                    N.ptr(gamma), N.ptr(beta), N.ptr(hmask), N.ptr(dy), N.ptr(dW1), N.ptr(db1),
                                    N.ptr(dW2), N.ptr(db2), N.ptr(dgamma), N.ptr(dbeta), N.ptr(dx), N.ptr(ws), rows,
                                    D, A, act, post, alpha, eps, N.stream()), "vlmclip_adapter_bwd")
        if need_dx:
            if x.shape != dx.shape or x.dtype != f32:
                # strided / bf16 views are only used on the frozen-backbone path, which never asks for dx
                raise N.NativeError("adapter dx is only available for contiguous fp32 inputs")
        return dx, dW1, db1, dW2, db2, dgamma, dbeta, None, None, None, None, None, None, None


def adapter(x, W1, b1, W2, b2, gamma=None, beta=None, *, act, post, alpha=0.0, eps=1e-5, hmask=None, ldx=None,
            rows=None):
    """y[rows, D] = post(act(x W1^T + b1) W2^T + b2, x).  `x` may be a wider buffer read with row stride `ldx`."""
    D = W1.shape[1]
    if rows is None:
        _req(x.dim() == 2 and x.shape[1] == D and x.stride(1) == 1, "adapter: x must be [R, D]")
        rows, ldx = x.shape[0], x.stride(0)
    for t in (W1, b1, W2, b2):
        _req(t.dtype == f32 and t.is_contiguous(), "adapter: weights must be contiguous fp32")
    return _AdapterFn.apply(x, W1, b1, W2, b2, gamma, beta, hmask, int(act), int(post), alpha, eps, int(ldx),
                            int(rows))


class _LinearF32Fn(torch.autograd.Function):
    """y = x W^T (+ b) with a FROZEN weight: backward only produces dx (projection heads, HF:784-785)."""

    @staticmethod
    def forward(ctx, x, W, b):
        R, K = x.shape
        Nn = W.shape[0]
        y = torch.empty((R, Nn), device=x.device, dtype=f32)
        N.check(N.load().vlmclip_linear_f32(N.ptr(x), x.stride(0), N.ptr(W), N.ptr(b), N.ptr(y), R, Nn, K, N.stream()),
                "vlmclip_linear_f32")
        ctx.save_for_backward(W)
        ctx.dims = (R, Nn, K)
        return y

    @staticmethod
    def backward(ctx, dy):
        (W,) = ctx.saved_tensors
        R, Nn, K = ctx.dims
        dx = None
        if ctx.needs_input_grad[0]:
            dy = dy.contiguous()
            dx = torch.empty((R, K), device=dy.device, dtype=f32)
            N.check(N.load().vlmclip_linear_f32_dgrad(N.ptr(dy), N.ptr(W), N.ptr(dx), R, Nn, K, N.stream()),
                    "vlmclip_linear_f32_dgrad")
        return dx, None, None


def linear_f32(x, W, b=None):
    _req(x.dtype == f32 and x.dim() == 2 and x.stride(1) == 1, "linear_f32: x must be fp32 [R, K]")
    _req(W.dtype == f32 and W.is_contiguous() and W.shape[1] == x.shape[1], "linear_f32: W must be fp32 [N, K]")
    return _LinearF32Fn.apply(x, W, b)


class _L2NormFn(torch.autograd.Function):
    """y = s/|s|, s = x (+ x2)."""

    @staticmethod
    def forward(ctx, x, x2):
        R, P = x.shape
        y = torch.empty_like(x)
        s = torch.empty_like(x) if x2 is not None else None
        N.check(N.load().vlmclip_l2norm_rows(N.ptr(x), N.ptr(x2), N.ptr(y), N.ptr(s), R, P, N.stream()),
                "vlmclip_l2norm_rows")
        ctx.save_for_backward(x if s is None else s)
        ctx.two = x2 is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        (s,) = ctx.saved_tensors
        R, P = s.shape
        dx = torch.empty_like(s)
        N.check(N.load().vlmclip_l2norm_rows_bwd(N.ptr(s), N.ptr(dy.contiguous()), N.ptr(dx), R, P, N.stream()),
                "vlmclip_l2norm_rows_bwd")
        return dx, (dx if ctx.two else None)


def l2norm(x, x2=None):
    """x / |x|, or (x + x2) / |x + x2| (average fusion followed by re-normalisation, model_v.py:306-315)."""
    _req(x.dtype == f32 and x.dim() == 2 and x.is_contiguous(), "l2norm: x must be contiguous fp32 [R, P]")
    if x2 is not None:
        _req(x2.dtype == f32 and x2.shape == x.shape and x2.is_contiguous(), "l2norm: x2 must match x")
    return _L2NormFn.apply(x, x2)


class _ScaledSimFn(torch.autograd.Function):
    """logits[B,C] = scale * f_img f_txt^T, differentiable (for callers applying their own criterion)."""

    @staticmethod
    def forward(ctx, f_img, f_txt, scale):
        B, P = f_img.shape
        Cc = f_txt.shape[0]
        logits = torch.empty((B, Cc), device=f_img.device, dtype=f32)
        N.check(
            N.load().vlmclip_class_head(N.ptr(f_img), N.ptr(f_txt), float(scale), None, None, N.ptr(logits), None, None,
                                        None, None, None, B, Cc, P, 1, N.stream()), "vlmclip_class_head")
        ctx.save_for_backward(f_img, f_txt)
        ctx.scale = float(scale)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        f_img, f_txt = ctx.saved_tensors
        B, P = f_img.shape
        Cc = f_txt.shape[0]
        d_img = torch.empty_like(f_img) if ctx.needs_input_grad[0] else None
        d_txt = torch.empty_like(f_txt) if ctx.needs_input_grad[1] else None
        if d_img is not None or d_txt is not None:
            N.check(
                N.load().vlmclip_class_head_bwd(N.ptr(f_img), N.ptr(f_txt), N.ptr(dlogits.contiguous()), ctx.scale,
                                                N.ptr(d_img), N.ptr(d_txt), B, Cc, P, N.stream()),
                "vlmclip_class_head_bwd")
        return d_img, d_txt, None


def scaled_similarity(f_img, f_txt, scale: float):
    for t in (f_img, f_txt):
        _req(t.dtype == f32 and t.dim() == 2 and t.is_contiguous(), "scaled_similarity: features must be contiguous fp32")
    return _ScaledSimFn.apply(f_img, f_txt, float(scale))


# --------------------------------------------------------------------------------------------- heads
_LOSS_COUNTERS = {}


def _loss_counters(dev, n_words: int):
    """Zeroed int32 words the strip kernels use for their last-arriver reductions (every launch leaves them zero), one
    buffer per (device, stream): launches on one stream are ordered, launches on different streams must not share it."""
    key = (str(dev), torch.cuda.current_stream(dev).cuda_stream)
    buf = _LOSS_COUNTERS.get(key)
    if buf is None or buf.numel() < n_words:
        buf = _LOSS_COUNTERS[key] = torch.zeros(max(1024, n_words), device=dev, dtype=torch.int32)
    return buf


class _ClipLossFn(torch.autograd.Function):
    """Symmetric InfoNCE over the (global) batch on strips of the logit matrix (csrc/clip_loss.cu); gradients for rows
    [row0, row0 + nloc) only.

    exchange=None, nloc == N   single process: 3 launches.
    exchange=None, nloc <  N   one process stands in for a rank and has no peer to ask: the strips of ALL rows are
                               computed (every LSE is then local) and only the local rows are differentiated.
    exchange=group             data parallel: strips of the local rows only; the ranks all-gather their
                               [lse_t | lse_i | loss share] blocks (2 nloc + 1 floats) between the two kernels.
    """

    @staticmethod
    def forward(ctx, txt_local, img_local, txt_all, img_all, scale_exp, row0, want_logits, logit_scale=None, exchange=None):
        Ng, P = txt_all.shape
        nloc = txt_local.shape[0]
        dev = txt_all.device
        lib = N.load()
        srow0, srows = (row0, nloc) if (exchange is not None or nloc == Ng) else (0, Ng)
        txt_n = torch.empty_like(txt_all)
        img_n = torch.empty_like(img_all)
        logits = torch.empty((Ng, Ng), device=dev, dtype=f32) if (want_logits and srows == Ng) else None
        block = torch.empty((2 * srows + 1,), device=dev, dtype=f32)  # [lse_t | lse_i | loss share]
        state = torch.empty(lib.vlmclip_clip_loss_state_size(Ng, P, srows), device=dev, dtype=f32)
        counters = _loss_counters(dev, lib.vlmclip_clip_loss_counters(srows))
        N.check(
            lib.vlmclip_clip_loss_fwd(N.ptr(txt_all), N.ptr(img_all), float(scale_exp), N.ptr(txt_n), N.ptr(img_n),
                                      N.ptr(logits), N.ptr(block), C_ptr_off(block, 2 * srows), N.ptr(state),
                                      N.ptr(counters), Ng, P, int(srow0), int(srows), N.stream()), "vlmclip_clip_loss_fwd")
        if exchange is not None:
            import torch.distributed as dist

            ws = dist.get_world_size(exchange if exchange is not True else None)
            gathered = torch.empty((ws, 2 * nloc + 1), device=dev, dtype=f32)
            dist.all_gather_into_tensor(gathered, block.view(1, -1), group=None if exchange is True else exchange)
            loss = gathered[:, 2 * nloc].sum()  # the one reduction left to torch: `ws` floats
            lse_all, lse_stride, rows_per_rank = gathered, 2 * nloc + 1, nloc
        else:
            loss = block[2 * srows]
            lse_all, lse_stride, rows_per_rank = block, 2 * srows + 1, srows
        # (autograd.Function.forward runs under no_grad: whether a gradient is wanted is read off the inputs)
        need_grad = txt_local.requires_grad or img_local.requires_grad or (logit_scale is not None)
        d_txt = d_img = d_ls = None
        if need_grad:
            d_txt = torch.empty((nloc, P), device=dev, dtype=f32)
            d_img = torch.empty((nloc, P), device=dev, dtype=f32)
            # trainable logit_scale (full fine-tune): also dL/d(log scale) for this rank's rows
            d_ls = torch.empty((1,), device=dev, dtype=f32) if logit_scale is not None else None
            ws_b = torch.empty(lib.vlmclip_clip_loss_bwd_workspace(Ng, P, nloc), device=dev, dtype=f32)
            N.check(
                lib.vlmclip_clip_loss_bwd(N.ptr(txt_n), N.ptr(img_n), N.ptr(lse_all), int(lse_stride), int(rows_per_rank),
                                          float(scale_exp), N.ptr(d_txt), N.ptr(d_img), N.ptr(d_ls), N.ptr(state),
                                          N.ptr(counters), N.ptr(ws_b), Ng, P, int(row0), nloc, int(srow0), int(srows),
                                          N.stream()), "vlmclip_clip_loss_bwd")
            ctx.save_for_backward(d_txt, d_img)
        ctx.has_grad = need_grad
        ctx.d_ls = d_ls
        ctx.ls_shape = tuple(logit_scale.shape) if logit_scale is not None else None
        ctx.mark_non_differentiable(txt_n, img_n)
        if logits is not None:
            ctx.mark_non_differentiable(logits)
            return loss, txt_n, img_n, logits
        return loss, txt_n, img_n, torch.empty(0, device=dev)

    @staticmethod
    def backward(ctx, dloss, *_):
        if not ctx.has_grad:
            return (None,) * 9
        d_txt, d_img = ctx.saved_tensors
        d_ls = None
        if ctx.d_ls is not None:
            d_ls = _scale_by(ctx.d_ls, dloss).reshape(ctx.ls_shape)
        return _scale_by(d_txt, dloss), _scale_by(d_img, dloss), None, None, None, None, None, d_ls, None


def _scale_by(g, dloss):
    """g * dloss with dloss a 0-d tensor (the upstream gradient of the scalar loss): on the library's kernel, not an
    aten elementwise op (vlmclip_scale_f32 broadcasts a device scalar)."""
    out = torch.empty_like(g)
    N.check(N.load().vlmclip_scale_f32(N.ptr(g), N.ptr(dloss.reshape(1).contiguous()), N.ptr(out), g.numel(), N.stream()),
            "vlmclip_scale_f32")
    return out


def C_ptr_off(t, off: int):
    """device pointer of element `off` of a contiguous tensor"""
    import ctypes

    return ctypes.c_void_p(t.data_ptr() + off * t.element_size())


def clip_loss(txt_local, img_local, logit_scale_exp: float, txt_all=None, img_all=None, row0: int = 0,
              want_logits: bool = True, logit_scale=None, exchange=None):
    """Returns (loss, txt_normalised_all, img_normalised_all, logits_per_text_all).

    txt_all / img_all: the all-gathered un-normalised features under data parallelism (default: the local ones).
    logit_scale: the (log) scale PARAMETER when it is trainable (full fine-tune); it then receives its gradient.
    exchange: a process group (or True for the default group) whose ranks each own `txt_local.shape[0]` consecutive rows:
    the log-sum-exps are then exchanged instead of being recomputed by every rank (see _ClipLossFn).  The full logit
    matrix is returned only when this process computes all of it (no exchange); otherwise an empty tensor.
    """
    if txt_all is None:
        txt_all, img_all, row0 = txt_local.detach(), img_local.detach(), 0
    for t in (txt_local, img_local, txt_all, img_all):
        _req(t.dtype == f32 and t.dim() == 2 and t.is_contiguous(), "clip_loss: features must be contiguous fp32 [N, P]")
    _req(txt_all.shape[1] % 4 == 0, "clip_loss: the projection dim must be a multiple of 4")
    return _ClipLossFn.apply(txt_local, img_local, txt_all, img_all, float(logit_scale_exp), int(row0), want_logits,
                             logit_scale, exchange)


class _ClassHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f_img, f_txt, scale, labels, soft):
        B, P = f_img.shape
        Cc = f_txt.shape[0]
        dev = f_img.device
        logits = torch.empty((B, Cc), device=dev, dtype=f32)
        have_lab = labels is not None or soft is not None
        loss = torch.zeros((1,), device=dev, dtype=f32)
        d_img = torch.empty_like(f_img) if have_lab else None
        d_txt = torch.empty_like(f_txt) if have_lab else None
        ws = torch.empty(B * Cc + B, device=dev, dtype=f32) if have_lab else None
        N.check(
            N.load().vlmclip_class_head(N.ptr(f_img), N.ptr(f_txt), float(scale), N.ptr(labels), N.ptr(soft),
                                        N.ptr(logits), None, N.ptr(loss) if have_lab else None, N.ptr(d_img),
                                        N.ptr(d_txt), N.ptr(ws), B, Cc, P, 1, N.stream()), "vlmclip_class_head")
        if have_lab:
            ctx.save_for_backward(d_img, d_txt)
        ctx.have_lab = have_lab
        ctx.mark_non_differentiable(logits)
        return loss[0], logits

    @staticmethod
    def backward(ctx, dloss, _dlogits):
        if not ctx.have_lab:
            raise N.NativeError("class_head: backward needs labels")
        d_img, d_txt = ctx.saved_tensors
        return _scale_by(d_img, dloss), _scale_by(d_txt, dloss), None, None, None


def class_head_loss(f_img, f_txt, scale: float, labels=None, soft_labels=None):
    """(loss, logits) of model_t.py:184-187: logits = scale * f_img f_txt^T, CE(logits, labels)."""
    for t in (f_img, f_txt):
        _req(t.dtype == f32 and t.dim() == 2 and t.is_contiguous(), "class_head: features must be contiguous fp32")
    if labels is not None:
        _req(labels.dtype == torch.int64 and labels.is_contiguous(), "class_head: labels must be int64")
    if soft_labels is not None:
        _req(soft_labels.dtype == f32 and soft_labels.is_contiguous(), "class_head: soft labels must be fp32")
    return _ClassHeadFn.apply(f_img, f_txt, float(scale), labels, soft_labels)


def class_head_probs(f_img, f_txt, scale: float, group: int = 1):
    """softmax(scale * f_img f_txt^T) with an optional max over `group` prompts per class (forward only)."""
    B, P = f_img.shape
    Cc = f_txt.shape[0] // group
    logits = torch.empty((B, Cc), device=f_img.device, dtype=f32)
    probs = torch.empty((B, Cc), device=f_img.device, dtype=f32)
    N.check(
        N.load().vlmclip_class_head(N.ptr(f_img.contiguous()), N.ptr(f_txt.contiguous()), float(scale), None, None,
                                    N.ptr(logits), N.ptr(probs), None, None, None, None, B, Cc, P, int(group),
                                    N.stream()), "vlmclip_class_head")
    return probs, logits


# --------------------------------------------------------------------------------------------- optimiser
# FusedAdamW updates parameters through raw pointers into its arena, behind autograd's version counters: anything
# that caches a derived copy of a trainable parameter (the bf16 weight packs of the inference paths) keys the cache on
# this generation as well as on Parameter._version.
_PARAM_GENERATION = 0


def param_generation() -> int:
    return _PARAM_GENERATION


class FusedAdamW:
    """clip_grad_norm_ + AdamW over one flat fp32 arena (trainer.py:91-99), two launches, no host sync.

    Parameters are re-pointed into a flat buffer (param.data becomes a view), and so are their .grad tensors,
    so autograd accumulates straight into the arena the kernel reads.
    """

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, max_grad_norm=0.0):
        self.params = [p for p in params]
        _req(len(self.params) > 0, "FusedAdamW: empty parameter list")
        dev = self.params[0].device
        # every parameter starts on a 16-byte boundary of the arena (the kernels read weights / biases with 16-byte
        # vector loads; a scalar such as logit_scale would otherwise shift everything behind it).  The gaps hold
        # zero parameters with zero gradients: they add nothing to the gradient norm and stay zero under AdamW.
        self.offsets = []
        n = 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.n = n
        self.flat = torch.zeros(n, device=dev, dtype=f32)
        self.grad = torch.zeros(n, device=dev, dtype=f32)
        self.exp_avg = torch.zeros(n, device=dev, dtype=f32)
        self.exp_avg_sq = torch.zeros(n, device=dev, dtype=f32)
        for p, off in zip(self.params, self.offsets):
            _req(p.dtype == f32 and p.device == dev, "FusedAdamW: parameters must be fp32 on one device")
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view(p.shape)
            p.grad = self.grad[off:off + k].view(p.shape)
        self.lr = torch.full((1,), float(lr), device=dev, dtype=f32)
        self._lr_on_device = float(lr)
        self.betas, self.eps, self.weight_decay, self.max_grad_norm = betas, eps, weight_decay, max_grad_norm
        self.step_t = torch.zeros((1,), device=dev, dtype=torch.int32)
        self.grad_norm = torch.zeros((1,), device=dev, dtype=f32)
        self.ws = torch.empty(1024, device=dev, dtype=f32)
        # torch.optim-compatible surface used by LR schedulers
        self.param_groups = [{"params": self.params, "lr": float(lr), "initial_lr": float(lr)}]

    def zero_grad(self, set_to_none: bool = False):
        self.grad.zero_()
        for p, off in zip(self.params, self.offsets):  # keep .grad pointing into the arena even if someone set it to None
            k = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * off:
                p.grad = self.grad[off:off + k].view(p.shape)

    def set_lr(self, lr: float):
        self.param_groups[0]["lr"] = float(lr)

    def push_lr(self):
        """Copy the host-side learning rate (an LR scheduler writes param_groups[0]["lr"], trainer.py:58-62,99) into the
        device scalar the kernel reads.  Called by step(); a CUDA-graph replay of step() calls it BEFORE the replay, the
        captured step itself must not contain the fill (it would bake one value into the graph)."""
        lr_host = float(self.param_groups[0]["lr"])
        if lr_host != self._lr_on_device:
            self.lr.fill_(lr_host)
            self._lr_on_device = lr_host

    def step(self):
        if not torch.cuda.is_current_stream_capturing():
            self.push_lr()
        N.check(
            N.load().vlmclip_adamw_clip_step(N.ptr(self.flat), N.ptr(self.grad), N.ptr(self.exp_avg),
                                             N.ptr(self.exp_avg_sq), self.n, N.ptr(self.lr), self.betas[0],
                                             self.betas[1], self.eps, self.weight_decay, self.max_grad_norm,
                                             N.ptr(self.step_t), N.ptr(self.grad_norm), N.ptr(self.ws), N.stream()),
            "vlmclip_adamw_clip_step")
        global _PARAM_GENERATION
        _PARAM_GENERATION += 1  # the parameters changed without their _version moving (see param_generation)

    def broadcast_from(self, src: int = 0, group=None):
        """Make every rank start from rank `src`'s parameters and optimiser state (what DDP's constructor does for the
        parameters): replicas that were built with different RNG state, or where only one rank loaded a checkpoint,
        would otherwise apply identical gradients to different weights."""
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        for t in (self.flat, self.exp_avg, self.exp_avg_sq, self.step_t):
            dist.broadcast(t, src=src, group=group)
        global _PARAM_GENERATION
        _PARAM_GENERATION += 1

    def state_dict(self):
        return {"params": self.flat.clone(), "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "step": self.step_t.clone(), "lr": self.param_groups[0]["lr"]}

    def load_state_dict(self, sd):
        global _PARAM_GENERATION
        if "params" in sd:
            _req(sd["params"].numel() == self.n, "FusedAdamW.load_state_dict: parameter arena size mismatch")
            self.flat.copy_(sd["params"])
            _PARAM_GENERATION += 1
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.step_t.copy_(sd["step"])
        self.set_lr(sd["lr"])
