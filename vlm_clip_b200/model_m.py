"""B200-native mirror of the reference's Track-M model (reference: model_m.py).

`CLIPWithAdapters` keeps the reference's constructor, attributes (`clip`, `processor`, `text_adapter`,
`vision_adapter`, `shared_adapters`), methods and return dicts, so `trainer.py` / `train.py` / `utest.py` work
against it unchanged.  What changes is who does the arithmetic:

  frozen towers     vlm_clip_b200.towers.NativeClipTowers (tcgen05 GEMMs, flash attention, folded LayerNorm)
  adapters          one fused kernel per adapter, evaluated on token 0 only — the reference applies the adapter
                    to all tokens and then keeps `[:, 0, :]` (model_m.py:87-90,102,116-118,122); the adapter is
                    position-wise, so this is result-identical
  projections       fp32 linear with an input-gradient kernel (weights frozen)
  loss              fused L2-normalise + scaled similarity + symmetric cross-entropy + gradient (model_m.py:146-171)

Reference quirks reproduced on purpose: text pooling takes token 0 (BOS), not EOS; the vision
adapter sees the states BEFORE post_layernorm and post_layernorm is never applied; `forward` returns normalised
features with the loss and un-normalised ones without it.

Data parallelism (new): after `enable_data_parallel()`, each rank all-gathers the un-normalised
features, evaluates the global N x N loss and differentiates only its own rows (exact dL_global/d local).
"""
from __future__ import annotations

import os
import warnings

import torch
import torch.nn as nn

from . import _native as N
from . import ops
from .adapter.clip_adapter import SharedMHSAttentionAdapter, TextAdapter, VisionAdapter
from .adapter.peclip import ContextAdapter, TextualAdapter
from .constants import CLIP_MEAN, CLIP_STD
from .finetune import TrainableClipTowers, linear_f32_trainable
from .towers import NativeClipTowers


def _load_clip(name):
    from transformers.models.clip import CLIPModel

    return CLIPModel.from_pretrained(name)


def _load_processor(name):
    from transformers.models.clip import CLIPProcessor

    try:
        return CLIPProcessor.from_pretrained(name)
    except Exception as e:  # offline box: the processor is host-side preprocessing, not part of the hot path
        warnings.warn(f"CLIPProcessor.from_pretrained({name!r}) failed ({type(e).__name__}); .processor is None")
        return None


class CLIPWithAdapters(nn.Module):
    """CLIP model with text and vision adapters (reference: model_m.py:10-248)."""

    def __init__(
        self,
        clip_model_name="openai/clip-vit-base-patch32",
        text_adapter_size=256,
        vision_adapter_size=256,
        shared_adapter_layers=2,
        freeze_clip=True,
        use_text_adapter=True,
        use_vision_adapter=True,
        use_shared_adapters=True,
        *,
        clip=None,
        processor=None,
        adapter_kind="clip_adapter",
    ):
        super().__init__()
        # `clip=` / `processor=` are extensions: a ready CLIPModel (e.g. random-init on an offline box)
        self.clip = clip if clip is not None else _load_clip(clip_model_name)
        self.processor = processor if processor is not None else (None if clip is not None else _load_processor(clip_model_name))

        text_hidden_size = self.clip.text_model.config.hidden_size
        vision_hidden_size = self.clip.vision_model.config.hidden_size

        self.use_text_adapter = use_text_adapter
        self.use_vision_adapter = use_vision_adapter
        self.use_shared_adapters = use_shared_adapters

        self.text_adapter = None
        self.vision_adapter = None
        self.shared_adapters = None
        # `adapter_kind` is an extension (BASELINE config 3): "peclip" puts adapter/peclip.py's TextualAdapter
        # (up(gelu(down x)) + x, no LayerNorm) into the same two slots; the reference wires PE-CLIP adapters into no model.
        # "peclip_context" puts peclip.ContextAdapter (LayerNorm(MHSA(x) + x) over the image patches, adapter/peclip.py:21-34)
        # into the vision slot and TextualAdapter into the text slot: the PE-CLIP pairing of a textual bottleneck with a
        # spatial-context adapter.  It mixes tokens, so it runs on the whole [B, S, D] stream before the CLS row is taken.
        if adapter_kind not in ("clip_adapter", "peclip", "peclip_context"):
            raise ValueError(f"adapter_kind must be 'clip_adapter', 'peclip' or 'peclip_context', got {adapter_kind!r}")
        self.adapter_kind = adapter_kind
        if self.use_text_adapter:
            self.text_adapter = (TextAdapter(text_hidden_size, text_adapter_size) if adapter_kind == "clip_adapter"
                                 else TextualAdapter(text_hidden_size, text_adapter_size))
        if self.use_vision_adapter:
            if adapter_kind == "clip_adapter":
                self.vision_adapter = VisionAdapter(vision_hidden_size, vision_adapter_size)
            elif adapter_kind == "peclip":
                self.vision_adapter = TextualAdapter(vision_hidden_size, vision_adapter_size)
            else:
                self.vision_adapter = ContextAdapter(vision_hidden_size, self.clip.vision_model.config.num_attention_heads)
        if self.use_shared_adapters:
            self.shared_adapters = nn.ModuleList(
                [SharedMHSAttentionAdapter(text_hidden_size, vision_hidden_size) for _ in range(shared_adapter_layers)])

        if freeze_clip:
            self._freeze_clip_parameters()

        self._towers = None
        self._towers_key = None
        self._ft_towers = None
        self._dp_group = None
        self._dp_enabled = False
        self.dp_return_logits = False  # data parallel: also return the full N x N logits_per_text (see _loss_outputs)
        # run the two towers on two CUDA streams (VLMCLIP_OVERLAP_TOWERS=0 serialises them, e.g. for per-kernel timing)
        self.overlap_towers = os.environ.get("VLMCLIP_OVERLAP_TOWERS", "1") != "0"
        self._tower_streams = None
        # Opt-in, result-preserving shortcut: Track M pools the text tower at token 0 (model_m.py:102, the
        # reference's BOS quirk) and the text tower is causal, so the pooled state depends on token 0 alone; with this
        # flag the text tower runs on that one token per caption.  Off by default: the dense tower is the benchmark.
        self.text_token0_only = False
        # Same kind of opt-in shortcut for the image side: the last vision layer evaluated for the CLS row only
        # (model_m.py:122 keeps last_hidden_state[:, 0]); needs the LN-folded towers.
        self.vision_cls_only_last_layer = False
        # on-GPU preprocessing of uint8 frames (extension): CLIP's normalisation by default, RGB channel order
        self.pixel_mean, self.pixel_std, self.frames_bgr = CLIP_MEAN, CLIP_STD, False

    # ------------------------------------------------------------------ reference API
    def _freeze_clip_parameters(self):
        for param in self.clip.parameters():
            param.requires_grad = False

    def _unfreeze_clip_parameters(self):
        for param in self.clip.parameters():
            param.requires_grad = True

    # ------------------------------------------------------------------ native backbone
    def _backbone(self) -> NativeClipTowers:
        p = next(self.clip.parameters())
        if p.device.type != "cuda":
            raise N.NativeError("CLIPWithAdapters runs on CUDA only (sm_100a kernels, no CPU fallback): call .to('cuda')")
        key = (str(p.device), p.data_ptr(), p._version)
        if self._towers is None or self._towers_key != key:
            self._towers = NativeClipTowers(self.clip, p.device)
            self._towers_key = key
            torch.cuda.current_stream().synchronize()  # packed weights are read from the private tower streams
        return self._towers

    def _full_finetune(self) -> bool:
        """True when CLIP parameters are trainable (`freeze_clip=False` / `_unfreeze_clip_parameters`, model_m.py:22,
        72-75; BASELINE config 5): the towers then run from the live fp32 parameters and are differentiated
        (finetune.TrainableClipTowers) instead of from the packed, LayerNorm-folded frozen copy."""
        return any(q.requires_grad for q in self.clip.parameters())

    def _finetune_towers(self) -> TrainableClipTowers:
        p = next(self.clip.parameters())
        if p.device.type != "cuda":
            raise N.NativeError("CLIPWithAdapters runs on CUDA only (sm_100a kernels, no CPU fallback): call .to('cuda')")
        if self._ft_towers is None or self._ft_towers.clip is not self.clip:
            self._ft_towers = TrainableClipTowers(self.clip)
        return self._ft_towers

    def _ft_text_features(self, input_ids, attention_mask):
        ft = self._finetune_towers()
        tok0 = ft.text_tok0(input_ids, attention_mask)  # final_layer_norm(hidden)[:, 0], fp32 [B, Dt]
        if self.use_text_adapter:
            tok0 = self.text_adapter(tok0)
        if self.use_shared_adapters:
            table = self.clip.vision_model.embeddings.position_embedding.weight.unsqueeze(0)
            for shared_adapter in self.shared_adapters:
                tok0 = shared_adapter(tok0.unsqueeze(1), table).squeeze(1)
        return linear_f32_trainable(tok0.contiguous(), ft.text_projection)

    def _ft_image_features(self, pixel_values):
        ft = self._finetune_towers()
        if pixel_values.dtype == torch.uint8:
            raise N.NativeError("full fine-tuning takes float pixel_values [B, 3, H, W] (uint8 frames: frozen towers only)")
        T = None
        if pixel_values.dim() == 5:
            B, C, T, H, W = pixel_values.shape
            pixel_values = pixel_values.permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W)
        cls = ft.vision_cls(pixel_values)  # last_hidden_state[:, 0] (no post_layernorm), fp32 [B, Dv]
        if self.use_vision_adapter:
            cls = self.vision_adapter(cls)
        feats = linear_f32_trainable(cls.contiguous(), ft.visual_projection)
        return feats if T is None else ops.mean_pool(feats, T)

    def refresh_backbone(self):
        """Re-pack the frozen CLIP weights (call after loading new backbone weights in place)."""
        self._towers = None

    def enable_data_parallel(self, group=None, enabled: bool = True):
        """Global-batch contrastive loss over `group` (default world): all-gather of features, local-row gradients."""
        self._dp_group = group
        self._dp_enabled = enabled

    def get_text_features(self, input_ids, attention_mask):
        """Text features with adapter: fp32 [B, P] (reference: model_m.py:77-105)."""
        if self._full_finetune():
            return self._ft_text_features(input_ids, attention_mask)
        bb = self._backbone()
        input_ids, attention_mask = self._text_inputs(input_ids, attention_mask)
        hidden = bb.text_stream_pre_ln(input_ids, attention_mask)  # [B*S, Dt] stream (towers.Hidden)
        return self._text_head(bb, hidden, input_ids.shape[0], input_ids.shape[1])

    def _text_inputs(self, input_ids, attention_mask):
        if self.text_token0_only and input_ids.shape[1] > 1:
            input_ids = input_ids[:, :1].contiguous()
            attention_mask = None if attention_mask is None else attention_mask[:, :1].contiguous()
        return input_ids, attention_mask

    def _text_head(self, bb, hidden, B, S):
        return self._text_head_rows(bb, hidden.rows_f32(B, S * bb.Dt, bb.Dt))

    def _text_head_rows(self, bb, rows):
        """rows: fp32 [B, Dt] = token 0 of every caption's pre-final-LN state (both terms of the residual stream)."""
        # final_layer_norm on the rows that are consumed (token 0 of every caption), in fp32
        tok0 = ops.layernorm_f32(rows, bb.final_ln_w, bb.final_ln_b, bb.eps_t)
        if self.use_text_adapter:
            tok0 = self.text_adapter(tok0)
        if self.use_shared_adapters:
            # model_m.py:93-100: every shared adapter attends from the text states to the vision position table; each
            # text row does so on its own and only token 0 is kept (model_m.py:102), so token 0 alone is evaluated
            # (bf16 tensor-core path under no_grad in eval mode, fp32 trainable path with dropout + backward otherwise).
            table = self.clip.vision_model.embeddings.position_embedding.weight.unsqueeze(0)
            for shared_adapter in self.shared_adapters:
                tok0 = shared_adapter(tok0.unsqueeze(1), table).squeeze(1)
        return ops.linear_f32(tok0, bb.text_projection)

    def get_image_features(self, pixel_values):
        """Image features with adapter: fp32 [B, P] (reference: model_m.py:107-125).

        `pixel_values`: float [B, 3, H, W] as in the reference, or (extension) decoded uint8 frames [B, Hs, Ws, 3], which
        are resized / scaled / normalised on the GPU (`self.pixel_mean`, `self.pixel_std`, `self.frames_bgr`)."""
        if self._full_finetune():
            return self._ft_image_features(pixel_values)
        bb = self._backbone()
        hidden, n, seq = self._vision_hidden(bb, pixel_values)
        return self._image_head(bb, hidden, n, seq)

    def get_video_features(self, clips):
        """Clip features: fp32 [B, P] = get_image_features(frames).view(B, T, P).mean(1) (the reference
        stops at `process_video`, no model consumes its output).  `clips`: float [B, 3, T, H, W] (stacked
        `process_video` outputs, process_video.py:29) or decoded uint8 frames [B, T, Hs, Ws, 3]."""
        if clips.dim() != 5:
            raise ValueError("clips must be [B, 3, T, H, W] float or [B, T, Hs, Ws, 3] uint8")
        if self._full_finetune():
            return self._ft_image_features(clips)
        T = clips.shape[1] if clips.dtype == torch.uint8 else clips.shape[2]
        bb = self._backbone()
        hidden, n, seq = self._vision_hidden(bb, clips)
        return ops.mean_pool(self._image_head(bb, hidden, n, seq), T)

    def _vision_hidden(self, bb, pixel_values):
        """Tower output for any accepted pixel layout -> (towers.Hidden stream [n*seq, D], n images, seq rows per image kept)."""
        cls_only = self.vision_cls_only_last_layer and bb.fold_ln
        seq = 1 if cls_only else bb.Sv
        if pixel_values.dtype == torch.uint8:
            if pixel_values.dim() not in (4, 5) or pixel_values.shape[-1] != 3:
                raise ValueError("uint8 frames must be [B, Hs, Ws, 3] or [B, T, Hs, Ws, 3]")
            n = pixel_values.numel() // (pixel_values.shape[-3] * pixel_values.shape[-2] * 3)
            return bb.vision_stream_u8(pixel_values, self.pixel_mean, self.pixel_std, self.frames_bgr, cls_only), n, seq
        if pixel_values.dim() == 5:  # [B, 3, T, H, W] -> [B*T, 3, H, W] (a layout copy, as the reference's caller would do)
            B, C, T, H, W = pixel_values.shape
            pixel_values = pixel_values.permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W)
        return bb.vision_stream(pixel_values, cls_only), pixel_values.shape[0], seq

    def _vision_adapter_mixes_tokens(self) -> bool:
        return self.use_vision_adapter and isinstance(self.vision_adapter, ContextAdapter)

    def _image_rows(self, bb, hidden, n, seq):
        """What the trainable half reads from the vision tower: the fp32 CLS rows [n, Dv] (hi + lo of the residual stream),
        or the whole bf16 stream [n, S, Dv] when the vision adapter mixes tokens."""
        if self._vision_adapter_mixes_tokens():
            if seq != bb.Sv:
                raise N.NativeError("vision_cls_only_last_layer cannot be combined with a token-mixing vision adapter")
            return hidden.hi.view(n, seq, bb.Dv)
        return hidden.rows_f32(n, seq * bb.Dv, bb.Dv)

    def _image_head(self, bb, hidden, B, seq=None):
        seq = bb.Sv if seq is None else seq
        return self._image_head_rows(bb, self._image_rows(bb, hidden, B, seq))

    def _image_head_rows(self, bb, cls):
        """cls: fp32 [n, Dv] = the CLS rows (hi + lo of the residual stream).  The adapter is position-wise, so evaluating
        it on row 0 of every sequence is result-identical to the reference's all-token call followed by [:, 0, :]
        (model_m.py:116-122)."""
        if cls.dim() == 3:
            # token-mixing vision adapter (ContextAdapter): cls is the whole bf16 stream [n, S, Dv]; the reference's order
            # is adapter on all tokens, then [:, 0, :] (model_m.py:116-122)
            cls = self.vision_adapter(cls)[:, 0, :].contiguous()
        elif self.use_vision_adapter:
            cls = self.vision_adapter(cls)
        return ops.linear_f32(cls, bb.visual_projection)

    def forward(self, input_ids=None, attention_mask=None, pixel_values=None, return_loss=True, *, inputs_ready=None):
        """Same contract as the reference (model_m.py:127-176): 5-key dict with the loss, 2-key dict without."""
        both = input_ids is not None and attention_mask is not None and pixel_values is not None
        if both and pixel_values.is_cuda and not self._full_finetune():
            rows_t, rows_v = self.tower_pooled(input_ids, attention_mask, pixel_values, inputs_ready)
            text_features, image_features = self.features_from_pooled(rows_t, rows_v)
        else:
            if input_ids is not None and attention_mask is not None:
                text_features = self.get_text_features(input_ids, attention_mask)
            else:
                text_features = None
            if pixel_values is not None:
                image_features = (self.get_video_features(pixel_values) if pixel_values.dim() == 5
                                  else self.get_image_features(pixel_values))
            else:
                image_features = None

        if return_loss and text_features is not None and image_features is not None:
            return self._loss_outputs(text_features, image_features)
        return {"text_features": text_features, "image_features": image_features}

    def _loss_outputs(self, text_features, image_features):
        """model_m.py:146-171: the 5-key dict with the symmetric InfoNCE loss (global batch under data parallelism)."""
        scale = self._logit_scale_exp()
        ls = self.clip.logit_scale
        ls_param = ls if (ls.requires_grad and torch.is_grad_enabled()) else None
        txt_all = img_all = exchange = None
        row0 = 0
        if self._dp_enabled:
            from .dist import gather_features, world

            if world(self._dp_group)[0] > 1:
                txt_all, img_all, row0 = gather_features(text_features, image_features, self._dp_group)
                exchange = self._dp_group if self._dp_group is not None else True
        loss, t_n, i_n, logits_per_text = ops.clip_loss(text_features, image_features, scale, txt_all, img_all, row0,
                                                        logit_scale=ls_param, exchange=exchange)
        if exchange is not None and self.dp_return_logits:
            # under data parallelism a rank only forms its strips of the global logit matrix; the full [N, N] matrix the
            # reference's dict carries is computed on request (evaluation scripts), not on the training path
            logits_per_text = ops.scaled_similarity(t_n, i_n, scale)
        return {
            "loss": loss,
            "text_features": t_n,
            "image_features": i_n,
            "logits_per_text": logits_per_text,
            "logits_per_image": logits_per_text.t(),
        }

    # ------------------------------------------------------------------ the step in two halves (frozen towers | trainable rest)
    def tower_pooled(self, input_ids, attention_mask, pixel_values, inputs_ready=None):
        """Frozen half of forward(): both towers -> the only rows the rest of the model reads, in fp32:
        (token 0 of every caption BEFORE final_layer_norm [B, Dt], CLS row of every image / frame [n, Dv]).
        Independent of every trainable parameter, so it may run ahead of the previous step's optimizer update
        (trainer.CLIPAdapterTrainer captures the two halves as two CUDA graphs and overlaps them across steps)."""
        bb = self._backbone()
        if self.overlap_towers:
            t_hidden, v_hidden, n_img, v_seq = self._both_towers(input_ids, attention_mask, pixel_values, inputs_ready)
        else:
            ids, am = self._text_inputs(input_ids, attention_mask)
            t_hidden = bb.text_stream_pre_ln(ids, am)
            v_hidden, n_img, v_seq = self._vision_hidden(bb, pixel_values)
        S = 1 if (self.text_token0_only and input_ids.shape[1] > 1) else input_ids.shape[1]
        rows_t = t_hidden.rows_f32(input_ids.shape[0], S * bb.Dt, bb.Dt)
        rows_v = self._image_rows(bb, v_hidden, n_img, v_seq)
        return rows_t, rows_v

    def features_from_pooled(self, rows_t, rows_v):
        """Trainable half of forward() up to the features: final LN, adapters, projections (and the temporal mean-pool
        when rows_v holds T frames per caption)."""
        bb = self._backbone()
        text_features = self._text_head_rows(bb, rows_t)
        image_features = self._image_head_rows(bb, rows_v)
        if rows_v.shape[0] != rows_t.shape[0]:  # clips: temporal mean-pool of the per-frame features
            image_features = ops.mean_pool(image_features, rows_v.shape[0] // rows_t.shape[0])
        return text_features, image_features

    def loss_from_pooled(self, rows_t, rows_v):
        return self._loss_outputs(*self.features_from_pooled(rows_t, rows_v))

    def _both_towers(self, input_ids, attention_mask, pixel_values, inputs_ready=None):
        """Frozen towers on two private streams, adapters / projections on the caller's stream.

        The towers are independent of each other until the loss (reference: model_m.py:127-150 calls them back to
        back) and, being frozen, independent of the optimizer.  Their kernels are persistent grids of one CTA per SM,
        so a tower alone leaves SMs idle in the last wave of every launch; with two streams the other tower's next
        kernel takes those SMs.  `inputs_ready` (an extension: a CUDA event after which the inputs are complete, e.g.
        `DevicePrefetcher`'s copy event) lets the towers of step i+1 start while the caller's stream still runs the
        latency-bound adapter backward / AdamW of step i; without it they wait for the caller's stream.
        """
        bb = self._backbone()
        main = torch.cuda.current_stream()
        if self._tower_streams is None:
            dev = pixel_values.device
            # default priority on purpose: high-priority tower streams starve the caller's stream (measured 11.8 vs
            # 11.6 ms per step), a high-priority caller's stream changes nothing (tools/overlap_ab.py)
            self._tower_streams = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
        vis, txt = self._tower_streams
        for st in (vis, txt):
            if inputs_ready is not None:
                st.wait_event(inputs_ready)
            else:
                st.wait_stream(main)
        input_ids, attention_mask = self._text_inputs(input_ids, attention_mask)
        with torch.cuda.stream(txt):
            t_hidden = bb.text_stream_pre_ln(input_ids, attention_mask)
        with torch.cuda.stream(vis):
            v_hidden, n_img, v_seq = self._vision_hidden(bb, pixel_values)
        input_ids.record_stream(txt)
        if attention_mask is not None:
            attention_mask.record_stream(txt)
        pixel_values.record_stream(vis)
        main.wait_stream(txt)
        main.wait_stream(vis)
        t_hidden.record_stream(main)
        v_hidden.record_stream(main)
        return t_hidden, v_hidden, n_img, v_seq

    def _logit_scale_exp(self) -> float:
        # logit_scale is a frozen scalar parameter: cache exp() on the host, re-read only when it is modified
        ls = self.clip.logit_scale
        if ls.requires_grad:
            # trainable scale (full fine-tune): the fused optimiser updates it in place behind autograd's version
            # counter, so it is read back every step (one 4-byte D2H copy against a ~100 ms step)
            return float(ls.detach().exp().item())
        key = (ls.data_ptr(), ls._version)
        if getattr(self, "_ls_key", None) != key:
            self._ls_val = float(ls.detach().exp().item())
            self._ls_key = key
        return self._ls_val

    # ------------------------------------------------------------------ checkpoints (reference: model_m.py:178-248)
    def save_adapter_weights(self, save_path):
        adapter_state_dict = {}
        if self.use_text_adapter:
            adapter_state_dict["text_adapter"] = self.text_adapter.state_dict()
        if self.use_vision_adapter:
            adapter_state_dict["vision_adapter"] = self.vision_adapter.state_dict()
        if self.use_shared_adapters:
            adapter_state_dict["shared_adapters"] = self.shared_adapters.state_dict()
        if not adapter_state_dict:
            raise ValueError("No adapters enabled to save")
        os.makedirs(os.path.dirname(save_path), exist_ok=True)
        torch.save(adapter_state_dict, save_path)
        print(f"Adapter weights saved to {save_path}")
        print(f"Saved adapters: {list(adapter_state_dict.keys())}")

    def load_adapter_weights(self, load_path):
        if not os.path.exists(load_path):
            raise FileNotFoundError(f"No adapter weights found at {load_path}")
        adapter_state_dict = torch.load(load_path, map_location=next(self.parameters()).device)
        for key, flag, module in (("text_adapter", self.use_text_adapter, self.text_adapter),
                                  ("vision_adapter", self.use_vision_adapter, self.vision_adapter),
                                  ("shared_adapters", self.use_shared_adapters, self.shared_adapters)):
            pretty = {"text_adapter": "Text adapter", "vision_adapter": "Vision adapter",
                      "shared_adapters": "Shared adapter"}[key]
            if key in adapter_state_dict:
                if not flag:
                    raise ValueError(f"{pretty} weights found but {pretty.lower()} is not enabled")
                # copy_ into the existing storage: parameters may live in the fused optimiser's flat arena
                module.load_state_dict(adapter_state_dict[key])
            elif flag:
                raise ValueError(f"{pretty} is enabled but no weights found in checkpoint")
        print(f"Adapter weights loaded from {load_path}")
        print(f"Loaded adapters: {list(adapter_state_dict.keys())}")
