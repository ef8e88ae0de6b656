"""Frozen CLIP towers executed by the sm_100a library (the L1 layer of SURVEY.md §1).

`NativeClipTowers` takes a HuggingFace `CLIPModel` purely as a weight container, packs its weights once
(bf16 matrices for the tensor cores, fp32 vectors, q/k/v fused into one [3D, D] matrix, LayerNorm folded
into the following dense layer) and replays the arithmetic of

    CLIPVisionTransformer.forward   HF modeling_clip.py:667-691  (embeddings :202-218, pre_layrnorm :677)
    CLIPTextTransformer.forward     HF modeling_clip.py:531-589  (embeddings :234-258, causal+padding mask :546-551)
    CLIPEncoderLayer.forward        HF modeling_clip.py:363-384  (attention :300-336, MLP :347-351)

with one kernel per fused step:

    [QKV GEMM: LN1 folded, bias] -> attention -> [out-proj GEMM: bias + residual in place, emits LN2 statistics]
    [fc1 GEMM: LN2 folded, bias, quick_gelu] -> [fc2 GEMM: bias + residual in place, emits next layer's LN1 statistics]

LayerNorm folding:  LN(x) W^T + b = rstd * (x (g*W)^T - mean * c) + (beta W^T + b),  c_n = sum_k (g*W)[n,k],
so the dense layer reads the raw bf16 residual stream and its epilogue applies the row statistics; the
normalised activations are never written to HBM.  The backbone is frozen (model_m.py:64-70), so there is no
backward through these layers.
"""
from __future__ import annotations

from dataclasses import dataclass
import os
from typing import Dict, List, Optional

import torch

from . import _native as N
from . import ops

bf16, f32 = torch.bfloat16, torch.float32


@dataclass
class _Layer:
    qkv_w: torch.Tensor  # [3D, D] bf16, gamma1 folded
    qkv_b: torch.Tensor  # [3D] fp32: beta1 W^T + b
    qkv_c: torch.Tensor  # [3D] fp32: row sums of the folded bf16 weight
    out_w: torch.Tensor
    out_b: torch.Tensor
    fc1_w: torch.Tensor  # [F, D] bf16, gamma2 folded
    fc1_b: torch.Tensor
    fc1_c: torch.Tensor
    fc2_w: torch.Tensor
    fc2_b: torch.Tensor
    # unfolded copies for the explicit-LayerNorm path (fold_ln=False; used by tests to bound the fold's error)
    ln1_w: torch.Tensor
    ln1_b: torch.Tensor
    ln2_w: torch.Tensor
    ln2_b: torch.Tensor
    qkv_w_raw: Optional[torch.Tensor] = None
    qkv_b_raw: Optional[torch.Tensor] = None
    fc1_w_raw: Optional[torch.Tensor] = None
    fc1_b_raw: Optional[torch.Tensor] = None


def _fold(W: torch.Tensor, b: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor):
    """Fold LN's affine into a following Linear.  Done once at pack time in fp32 (weight preparation, not hot path)."""
    Wf = (W * gamma[None, :]).to(bf16)
    c = Wf.float().sum(dim=1).contiguous()
    d = (W @ beta + b).contiguous()
    return Wf.contiguous(), d, c


def _pack_layers(sd: Dict[str, torch.Tensor], prefix: str, dev, keep_raw: bool) -> List[_Layer]:
    layers = []
    l = 0
    while f"{prefix}encoder.layers.{l}.layer_norm1.weight" in sd:
        p = f"{prefix}encoder.layers.{l}."
        g = lambda k: sd[p + k].detach().to(dev, f32)
        Wqkv = torch.cat([g("self_attn.q_proj.weight"), g("self_attn.k_proj.weight"), g("self_attn.v_proj.weight")], 0)
        bqkv = torch.cat([g("self_attn.q_proj.bias"), g("self_attn.k_proj.bias"), g("self_attn.v_proj.bias")], 0)
        qkv_w, qkv_b, qkv_c = _fold(Wqkv, bqkv, g("layer_norm1.weight"), g("layer_norm1.bias"))
        fc1_w, fc1_b, fc1_c = _fold(g("mlp.fc1.weight"), g("mlp.fc1.bias"), g("layer_norm2.weight"), g("layer_norm2.bias"))
        layers.append(
            _Layer(qkv_w, qkv_b, qkv_c, g("self_attn.out_proj.weight").to(bf16).contiguous(),
                   g("self_attn.out_proj.bias").contiguous(), fc1_w, fc1_b, fc1_c,
                   g("mlp.fc2.weight").to(bf16).contiguous(), g("mlp.fc2.bias").contiguous(),
                   g("layer_norm1.weight").contiguous(), g("layer_norm1.bias").contiguous(),
                   g("layer_norm2.weight").contiguous(), g("layer_norm2.bias").contiguous(),
                   Wqkv.to(bf16).contiguous() if keep_raw else None, bqkv.contiguous() if keep_raw else None,
                   g("mlp.fc1.weight").to(bf16).contiguous() if keep_raw else None,
                   g("mlp.fc1.bias").contiguous() if keep_raw else None))
        l += 1
    return layers


_STATS_IN_GEMM = os.environ.get("VLMCLIP_STATS_IN_GEMM", "0") == "1"  # experiment switch
# Residual GEMMs finalise the LN statistics themselves (last-arriving CTA of a row block) instead of a separate
# ln_partials_to_stats launch.  OFF: interleaved A/B on B200 (tools/ab_step.py fused_stats, profiles/r02_ab_interleaved.txt)
# measured 13.31 ms per step fused against 12.94 ms with the 46 separate launches - the two extra barriers that join
# both epilogue groups at the end of every tile cost more than the launch-bound statistics kernels they replace.
_FUSED_STATS = os.environ.get("VLMCLIP_FUSED_STATS", "0") == "1"


class Hidden:
    """A tower's output stream [B*S, D].  `hi` is the bf16 tensor every bf16 consumer reads; with the two-term residual
    stream (`NativeClipTowers(residual="hilo")`, the default) `lo` holds bf16(x - hi) and fp32 consumers (the pooled
    rows that feed adapters / projections) read hi + lo."""

    __slots__ = ("hi", "lo")

    def __init__(self, hi: torch.Tensor, lo: Optional[torch.Tensor] = None):
        self.hi, self.lo = hi, lo

    def rows_f32(self, rows: int, ld: int, D: int) -> torch.Tensor:
        """fp32 [rows, D]: row r = elements [r*ld, r*ld + D) of the stream (token 0 of every sequence for ld = S*D)."""
        return ops.gather_rows_f32(self.hi, rows, ld, D, lo=self.lo)

    def record_stream(self, stream) -> None:
        self.hi.record_stream(stream)
        if self.lo is not None:
            self.lo.record_stream(stream)


class NativeClipTowers:
    """Device-resident packed weights + forward of both frozen towers."""

    def __init__(self, clip, device=None, fold_ln: bool = True, keep_raw: bool = False, residual: Optional[str] = None):
        """residual: "hilo" (default; VLMCLIP_RESIDUAL overrides) keeps the residual stream as two bf16 planes hi + lo
        (vlmclip_gemm_bf16_res2) so that its 2 x L in-place updates do not accumulate bf16 rounding; "bf16" keeps one
        plane (less HBM traffic, ~2.5x the end-to-end error: oracle/emulate_bf16.py).  LN folding is required for
        "hilo" (the unfolded path exists to bound the fold's error in tests)."""
        N.load()
        residual = residual or os.environ.get("VLMCLIP_RESIDUAL", "hilo")
        if residual not in ("hilo", "bf16"):
            raise ValueError(f"residual must be 'hilo' or 'bf16', got {residual!r}")
        self.residual = residual if fold_ln else "bf16"
        dev = torch.device(device) if device is not None else next(clip.parameters()).device
        if dev.type != "cuda":
            raise N.NativeError("NativeClipTowers needs a CUDA device: the towers only run on the sm_100a library")
        self.device = dev
        self.fold_ln = fold_ln
        self._tables = {}
        cfg = clip.config
        vc, tc = cfg.vision_config, cfg.text_config
        self.eps_v, self.eps_t = float(vc.layer_norm_eps), float(tc.layer_norm_eps)
        self.Dv, self.Dt = vc.hidden_size, tc.hidden_size
        self.Hv, self.Ht = vc.num_attention_heads, tc.num_attention_heads
        self.patch, self.image = vc.patch_size, vc.image_size
        self.Sv = (vc.image_size // vc.patch_size) ** 2 + 1
        self.vocab = tc.vocab_size
        self.max_pos = tc.max_position_embeddings
        if self.Dv // self.Hv != 64 or self.Dt // self.Ht != 64:
            raise ValueError("the attention kernel is specialised for head_dim = 64 (all OpenAI CLIP towers)")
        if vc.hidden_act != "quick_gelu" or tc.hidden_act != "quick_gelu":
            raise ValueError("only quick_gelu towers are supported (OpenAI CLIP); got %s/%s" % (vc.hidden_act, tc.hidden_act))
        sd = clip.state_dict()
        keep_raw = keep_raw or not fold_ln
        g = lambda k: sd[k].detach().to(dev, f32).contiguous()
        # ---- vision embeddings ----
        w = g("vision_model.embeddings.patch_embedding.weight").reshape(self.Dv, -1)
        K = w.shape[1]
        Kpad = (K + 63) // 64 * 64
        wp = torch.zeros(self.Dv, Kpad, device=dev, dtype=f32)
        wp[:, :K] = w
        self.patch_w = wp.to(bf16).contiguous()
        self.cls = g("vision_model.embeddings.class_embedding")
        self.pos_v = g("vision_model.embeddings.position_embedding.weight")
        self.pre_ln_w, self.pre_ln_b = g("vision_model.pre_layrnorm.weight"), g("vision_model.pre_layrnorm.bias")
        self.post_ln_w, self.post_ln_b = g("vision_model.post_layernorm.weight"), g("vision_model.post_layernorm.bias")
        self.v_layers = _pack_layers(sd, "vision_model.", dev, keep_raw)
        # ---- text embeddings ----
        self.tok = g("text_model.embeddings.token_embedding.weight")
        self.pos_t = g("text_model.embeddings.position_embedding.weight")
        self.final_ln_w, self.final_ln_b = g("text_model.final_layer_norm.weight"), g("text_model.final_layer_norm.bias")
        self.t_layers = _pack_layers(sd, "text_model.", dev, keep_raw)
        # ---- heads (frozen, fp32: the trainable path runs in fp32) ----
        self.visual_projection = g("visual_projection.weight")
        self.text_projection = g("text_projection.weight")
        self.eos_token_id = getattr(tc, "eos_token_id", 2)

    # ------------------------------------------------------------------------------------------ encoder
    def _layer_table(self, layers):
        """ctypes array of vlmclip_layer_t for `layers` (built once; the tensors it points at are owned by `layers`)."""
        key = id(layers)
        tab = self._tables.get(key)
        if tab is None:
            tab = (N.LayerPtrs * len(layers))()
            for i, L in enumerate(layers):
                for name in ("qkv_w", "qkv_b", "qkv_c", "out_w", "out_b", "fc1_w", "fc1_b", "fc1_c", "fc2_w", "fc2_b"):
                    t = getattr(L, name)
                    if not (t.is_cuda and t.is_contiguous()):
                        raise N.NativeError(f"layer {i}: {name} must be a contiguous CUDA tensor")
                    setattr(tab[i], name, t.data_ptr())
            self._tables[key] = tab
        return tab

    def _row_counters(self, layers, M: int, dev):
        """Zeroed int32 words for the fused LayerNorm statistics of the residual GEMMs (one per 128-row block; every
        launch leaves them zero).  One buffer per tower: the towers run concurrently on two streams.
        Returns None (separate ln_partials_to_stats launches) unless VLMCLIP_FUSED_STATS=1: see _FUSED_STATS."""
        if not _FUSED_STATS:
            return None
        key = ("cnt", id(layers))
        buf = self._tables.get(key)
        need = (M + 127) // 128
        if buf is None or buf.numel() < need or buf.device != dev:
            buf = self._tables[key] = torch.zeros(max(need, 1024), device=dev, dtype=torch.int32)
        return buf

    def _new_stream(self, M: int, D: int, dev) -> torch.Tensor:
        """Storage of a residual stream: bf16 [P, M, D] with P = 2 planes (hi, lo) for the two-term stream, else 1."""
        return torch.empty((2 if self.residual == "hilo" else 1, M, D), device=dev, dtype=bf16)

    def _encoder(self, x2: torch.Tensor, layers: List[_Layer], B: int, S: int, H: int, eps: float, causal: bool,
                 key_mask: Optional[torch.Tensor], cls_only: bool = False) -> Hidden:
        """All encoder layers over the residual stream x2 [P, B*S, D] (in place; see `_new_stream`).  `cls_only` (vision,
        LN folded): the last layer is evaluated for token 0 of every sequence only and [B, D] is returned (see
        `_last_layer_cls`)."""
        P, M, D = x2.shape
        x = x2[0]
        x_lo = x2[1] if P == 2 else None
        F = layers[0].fc1_w.shape[0]
        dev = x.device
        qkv = torch.empty((M, 3 * D), device=dev, dtype=bf16)
        att = torch.empty((M, D), device=dev, dtype=bf16)
        hid = torch.empty((M, F), device=dev, dtype=bf16)
        stats = torch.empty((M, 2), device=dev, dtype=f32)
        # per-32-column (mean, M2) partials of the residual stream, written by the epilogue of the layer that produces
        # x (out-proj / fc2); a 5 MB combine pass turns them into (mean, rstd) instead of re-reading the 77 MB stream
        part = torch.empty((M, D // 32, 2), device=dev, dtype=f32)
        xn = None if self.fold_ln else torch.empty((M, D), device=dev, dtype=bf16)
        if cls_only and not (self.fold_ln and not causal and key_mask is None):
            raise N.NativeError("cls_only needs the LN-folded, unmasked (vision) encoder")
        if self.fold_ln and (cls_only or (ops.TRACE is None and ops.PROFILE is None and not _STATS_IN_GEMM)):
            # the whole tower in one native call (same kernels, same order as the loop below: that loop stays as the
            # per-op path the tracing / roofline hooks time)
            n_full = len(layers) - 1 if cls_only else len(layers)
            if n_full > 0:
                table = self._layer_table(layers)
                N.check(
                    N.load().vlmclip_encoder_fwd(table, n_full, N.ptr(x), N.ptr(x_lo), N.ptr(qkv), N.ptr(att), N.ptr(hid),
                                                 N.ptr(stats), N.ptr(part), N.ptr(self._row_counters(layers, M, dev)),
                                                 N.ptr(key_mask), B, S, H, D, F, float(eps), 1 if causal else 0,
                                                 N.ACT_QUICK_GELU, N.stream()), "vlmclip_encoder_fwd")
            if cls_only:
                return Hidden(self._last_layer_cls(x, L=layers[-1], B=B, S=S, H=H, eps=eps, qkv=qkv, stats=stats, part=part,
                                                   first=n_full == 0))
            return Hidden(x, x_lo)

        def residual_gemm(a, w, b):
            # x += a w^T + b: the same thread reads and writes each element, so the update is done in place
            if x_lo is not None:
                ops.gemm_res2(a, w, b, x2, stats_part_out=part)
            else:
                ops.gemm(a, w, bias=b, residual=x, out=x, stats_part_out=part if self.fold_ln else None)

        first = True
        for L in layers:
            if self.fold_ln:
                if first:  # the embeddings kernel does not emit partials: one explicit statistics pass
                    ops.row_stats(x, eps, out=stats)
                    first = False
                    ops.gemm(x, L.qkv_w, bias=L.qkv_b, row_stats=stats, col_c=L.qkv_c, out=qkv)
                elif _STATS_IN_GEMM:
                    ops.gemm(x, L.qkv_w, bias=L.qkv_b, stats_part_in=part, ln_eps=eps, col_c=L.qkv_c, out=qkv)
                else:
                    ops.ln_partials_to_stats(part, eps, out=stats)
                    ops.gemm(x, L.qkv_w, bias=L.qkv_b, row_stats=stats, col_c=L.qkv_c, out=qkv)
            else:
                ops.layernorm(x, L.ln1_w, L.ln1_b, eps, out=xn)
                ops.gemm(xn, L.qkv_w_raw, bias=L.qkv_b_raw, out=qkv)
            ops.attention(qkv, B, S, H, causal=causal, key_mask=key_mask, out=att)
            residual_gemm(att, L.out_w, L.out_b)
            if self.fold_ln and _STATS_IN_GEMM:
                ops.gemm(x, L.fc1_w, bias=L.fc1_b, stats_part_in=part, ln_eps=eps, col_c=L.fc1_c, act=N.ACT_QUICK_GELU,
                         out=hid)
            elif self.fold_ln:
                ops.ln_partials_to_stats(part, eps, out=stats)
                ops.gemm(x, L.fc1_w, bias=L.fc1_b, row_stats=stats, col_c=L.fc1_c, act=N.ACT_QUICK_GELU, out=hid)
            else:
                ops.layernorm(x, L.ln2_w, L.ln2_b, eps, out=xn)
                ops.gemm(xn, L.fc1_w_raw, bias=L.fc1_b_raw, act=N.ACT_QUICK_GELU, out=hid)
            residual_gemm(hid, L.fc2_w, L.fc2_b)
        return Hidden(x, x_lo)

    def _last_layer_cls(self, x, L: _Layer, B: int, S: int, H: int, eps: float, qkv, stats, part, first: bool):
        """Last encoder layer for token 0 only -> bf16 [B, D] = last_hidden_state[:, 0].

        Result-preserving: Track M keeps `last_hidden_state[:, 0]` (model_m.py:122), and within a layer
        token 0 needs every token's K and V but only its own Q, attention row, out-proj, LN2 and MLP.  The QKV GEMM
        still runs on all tokens; everything after it runs on B rows."""
        D = x.shape[1]
        if first:
            ops.row_stats(x, eps, out=stats)
        else:
            ops.ln_partials_to_stats(part, eps, out=stats)
        ops.gemm(x, L.qkv_w, bias=L.qkv_b, row_stats=stats, col_c=L.qkv_c, out=qkv)
        q0 = qkv.view(B, S, 3 * D)[:, 0, :D]  # row stride S*3D
        att0 = ops.attention_1q(q0, qkv[:, D:2 * D], qkv[:, 2 * D:], S, H, kv_row_stride=3 * D, kv_batch_stride=S * 3 * D)
        x0 = x.view(B, S, D)[:, 0]  # row stride S*D
        y0 = ops.gemm(att0, L.out_w, bias=L.out_b, residual=x0)
        st0 = ops.row_stats(y0, eps)
        h0 = ops.gemm(y0, L.fc1_w, bias=L.fc1_b, row_stats=st0, col_c=L.fc1_c, act=N.ACT_QUICK_GELU)
        return ops.gemm(h0, L.fc2_w, bias=L.fc2_b, residual=y0, out=y0)

    # ------------------------------------------------------------------------------------------ towers
    @torch.no_grad()
    def vision_stream(self, pixel_values: torch.Tensor, cls_only: bool = False) -> Hidden:
        """`vision_model(pixel_values).last_hidden_state` (NOT through post_layernorm, HF:684) as a `Hidden` stream
        [B*S, D]; with `cls_only` just its token-0 rows, [B, D]."""
        if pixel_values.dim() != 4 or pixel_values.shape[1] != 3:
            raise ValueError("pixel_values must be [B, 3, H, W]")
        if pixel_values.shape[2] != self.image or pixel_values.shape[3] != self.image:
            raise ValueError(f"Input image size ({pixel_values.shape[2]}*{pixel_values.shape[3]}) doesn't match model "
                             f"({self.image}*{self.image}).")  # same check as HF:204-207
        if pixel_values.dtype not in (f32, bf16):
            pixel_values = pixel_values.float()
        pixel_values = pixel_values.contiguous()
        B = pixel_values.shape[0]
        return self._vision_from_cols(ops.im2col(pixel_values, self.patch), B, cls_only)

    def vision_hidden(self, pixel_values: torch.Tensor, cls_only: bool = False) -> torch.Tensor:
        """bf16 view of `vision_stream` (the hi plane)."""
        return self.vision_stream(pixel_values, cls_only).hi

    def _vision_from_cols(self, cols: torch.Tensor, B: int, cls_only: bool = False) -> Hidden:
        patches = ops.gemm(cols, self.patch_w)  # [B*np, D] bf16 (staged TMA epilogue; the fp32 path costs 2x the traffic)
        x2 = self._new_stream(B * self.Sv, self.Dv, cols.device)
        ops.vision_embed_ln(patches, self.cls, self.pos_v, self.pre_ln_w, self.pre_ln_b, B, self.Sv, self.eps_v, out=x2[0],
                            out_lo=x2[1] if x2.shape[0] == 2 else None)
        return self._encoder(x2, self.v_layers, B, self.Sv, self.Hv, self.eps_v, False, None, cls_only)

    @torch.no_grad()
    def vision_stream_u8(self, frames_u8: torch.Tensor, mean, std, bgr: bool = False, cls_only: bool = False) -> Hidden:
        """Same as `vision_stream`, from decoded uint8 frames [..., Hs, Ws, 3]: resize to the model resolution, /255 and
        normalisation are fused into the patch extraction (process_video.py:14-29 semantics; `vlmclip_preprocess_patches`)."""
        if frames_u8.dtype != torch.uint8 or frames_u8.dim() < 4 or frames_u8.shape[-1] != 3:
            raise ValueError("frames must be uint8 [..., Hs, Ws, 3]")
        frames_u8 = frames_u8.contiguous()
        n = frames_u8.numel() // (frames_u8.shape[-3] * frames_u8.shape[-2] * 3)
        cols = ops.preprocess_patches(frames_u8, self.image, self.image, self.patch, mean, std, bgr)
        return self._vision_from_cols(cols, n, cls_only)

    def vision_hidden_u8(self, frames_u8, mean, std, bgr: bool = False, cls_only: bool = False) -> torch.Tensor:
        return self.vision_stream_u8(frames_u8, mean, std, bgr, cls_only).hi

    @torch.no_grad()
    def text_stream_pre_ln(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor]) -> Hidden:
        """Text encoder output BEFORE final_layer_norm, `Hidden` stream [B*S, D]."""
        if input_ids.dim() != 2:
            raise ValueError("input_ids must be [B, S]")
        B, S = input_ids.shape
        if S > self.max_pos:
            raise ValueError(f"Sequence length must be less than max_position_embeddings (got `sequence length`: {S} "
                             f"and max_position_embeddings: {self.max_pos}")
        ids = input_ids.to(torch.int64).contiguous()
        key_mask = None
        if attention_mask is not None:
            key_mask = (attention_mask != 0).to(torch.uint8).contiguous()
        x2 = self._new_stream(B * S, self.Dt, ids.device)
        ops.text_embed(ids, self.tok, self.pos_t, out=x2[0], out_lo=x2[1] if x2.shape[0] == 2 else None)
        return self._encoder(x2, self.t_layers, B, S, self.Ht, self.eps_t, True, key_mask)

    def text_hidden_pre_ln(self, input_ids, attention_mask) -> torch.Tensor:
        return self.text_stream_pre_ln(input_ids, attention_mask).hi

    @torch.no_grad()
    def text_hidden(self, input_ids, attention_mask) -> torch.Tensor:
        """`text_model(...).last_hidden_state` (after final_layer_norm, HF:562) as bf16 [B*S, D]."""
        x = self.text_hidden_pre_ln(input_ids, attention_mask)
        return ops.layernorm(x, self.final_ln_w, self.final_ln_b, self.eps_t)

    # ------------------------------------------------------------------------------------------ pooled features
    @torch.no_grad()
    def image_features(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """CLIPModel.get_image_features (HF:829-863): projection(post_layernorm(CLS)), fp32 [B, P]."""
        h = self.vision_stream(pixel_values)
        B = pixel_values.shape[0]
        cls = h.rows_f32(B, self.Sv * self.Dv, self.Dv)
        pooled = ops.layernorm_f32(cls, self.post_ln_w, self.post_ln_b, self.eps_v)
        return ops.linear_f32(pooled, self.visual_projection)

    @torch.no_grad()
    def text_features(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor]) -> torch.Tensor:
        """CLIPModel.get_text_features (HF:793-827): projection of the final-LN state at the EOS position."""
        h = self.text_stream_pre_ln(input_ids, attention_mask)
        B, S = input_ids.shape
        ids = input_ids.to(torch.int)
        if self.eos_token_id == 2:
            eos = ids.argmax(dim=-1)  # HF:564-575 (index bookkeeping on [B,S] ints, not arithmetic on activations)
        else:
            eos = (ids == self.eos_token_id).int().argmax(dim=-1)
        rows = (torch.arange(B, device=h.hi.device) * S + eos).to(torch.int64)
        picked = Hidden(h.hi.index_select(0, rows).contiguous(),
                        None if h.lo is None else h.lo.index_select(0, rows).contiguous())
        pooled = ops.layernorm_f32(picked.rows_f32(B, self.Dt, self.Dt), self.final_ln_w, self.final_ln_b, self.eps_t)
        return ops.linear_f32(pooled, self.text_projection)
