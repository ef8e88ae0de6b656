"""B200-native mirror of the reference's Track-V model (reference: model_v.py).

`EnhancedCLIPAdapter` = Track T plus a third "context" branch fed by caption embeddings (gamma blend, average
fusion, re-normalise) and CLIP's learned temperature (model_v.py:260-343).  The caption generator
(`VLMContextExtractor`, 4-bit Qwen2.5-VL, model_v.py:43-142) is OUT OF SCOPE: pass the
context features in, exactly as `forward` already expects.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _native as N
from . import constants as config
from . import ops
from .model_m import _load_clip, _load_processor
from .model_t import _ReluBottleneck
from .towers import NativeClipTowers


class BaseAdapter(_ReluBottleneck):
    """fc2(dropout(relu(fc1(x)))) (reference: model_v.py:18-27); dropout is applied as a mask inside the fused kernel."""

    def __init__(self, input_dim, bottleneck_dim):
        super().__init__(input_dim, bottleneck_dim)
        self.dropout = nn.Dropout(0.1)

    def _mask(self, rows: int, dev):
        if not self.training or self.dropout.p == 0.0:
            return None
        keep = 1.0 - self.dropout.p
        A = self.fc1.out_features
        return (torch.rand(rows, A, device=dev) < keep).float() / keep

    def forward(self, x):
        x2 = x.reshape(-1, x.shape[-1]).float().contiguous()
        y = ops.adapter(x2, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, act=N.ACT_RELU,
                        post=N.POST_PLAIN, hmask=self._mask(x2.shape[0], x2.device))
        return y.view(*x.shape)

    def blend(self, x, ratio: float, hmask=None):
        x2 = x.reshape(-1, x.shape[-1]).float().contiguous()
        return super().blend(x2, ratio, hmask=self._mask(x2.shape[0], x2.device))


class ContextAdapter(BaseAdapter):
    pass


class VisualAdapter(BaseAdapter):
    pass


class TextAdapter(BaseAdapter):
    pass


class VLMContextExtractor:
    def __init__(self, *a, **k):
        raise N.NativeError("VLMContextExtractor (Qwen2.5-VL caption generation) is outside the accelerated hot path; "
                            "compute context features elsewhere and pass them to EnhancedCLIPAdapter.forward")


class EnhancedCLIPAdapter(nn.Module):
    def __init__(self, clip_model_name=config.CLIP_MODEL_NAME, alpha=config.ALPHA, beta=config.BETA, gamma=config.GAMMA,
                 bottleneck_dim=config.V_BOTTLENECK_DIM, device=config.DEVICE, vlm_context_extractor=None, *, clip=None,
                 processor=None):
        super().__init__()
        self.device = device
        self.model = (clip if clip is not None else _load_clip(clip_model_name)).to(self.device)
        self.processor = processor if processor is not None else (None if clip is not None else _load_processor(clip_model_name))
        for param in self.model.parameters():
            param.requires_grad = False
        self.image_feature_dim = self.model.config.projection_dim
        self.text_feature_dim = self.image_feature_dim
        self.visual_adapter = VisualAdapter(self.image_feature_dim, bottleneck_dim).to(self.device)
        self.text_adapter = TextAdapter(self.text_feature_dim, bottleneck_dim).to(self.device)
        self.context_adapter = ContextAdapter(self.text_feature_dim, bottleneck_dim).to(self.device)
        self.alpha, self.beta, self.gamma = alpha, beta, gamma
        self.vlm_context_extractor = vlm_context_extractor  # never built here (out of scope)
        self.original_emotion_text_features = {}
        self.adapted_emotion_embedding_tensor = None
        self.emotion_embedding_tensor = None
        self._towers = None

    def _backbone(self) -> NativeClipTowers:
        if self._towers is None:
            self._towers = NativeClipTowers(self.model, next(self.model.parameters()).device)
        return self._towers

    def encode_emotion_descriptions(self, emotions=config.EMOTIONS):
        """One prompt per class, "A person expressing {emotion}" (reference: model_v.py:196-238)."""
        if self.processor is None:
            raise N.NativeError("no CLIPProcessor available offline: set emotion_embedding_tensor directly")
        dev = next(self.model.parameters()).device
        self.emotion_descriptions = {e: [f"A person expressing {e}"] for e in emotions}
        self.original_emotion_text_features = {}
        for emotion, descriptions in self.emotion_descriptions.items():
            feats = []
            for description in descriptions:
                enc = self.processor(text=[description], padding=True, truncation=True, return_tensors="pt")
                feats.append(ops.l2norm(self._backbone().text_features(enc["input_ids"].to(dev), enc["attention_mask"].to(dev))))
            self.original_emotion_text_features[emotion] = torch.cat(feats, 0).mean(dim=0, keepdim=True)
        self.emotion_embedding_tensor = torch.cat(list(self.original_emotion_text_features.values()), dim=0).to(dev)
        self.update_emotion_embeddings()

    def update_emotion_embeddings(self):
        if self.emotion_embedding_tensor is None:
            print("Warning: Original emotion embeddings not encoded. Call encode_emotion_descriptions first.")
            return
        with torch.no_grad():
            was = self.text_adapter.training
            self.text_adapter.eval()  # the reference runs this under no_grad with the module's current mode; dropout
            self.adapted_emotion_embedding_tensor = self.text_adapter.blend(self.emotion_embedding_tensor, self.beta)
            self.text_adapter.train(was)

    def features(self, pixel_values, context_features=None):
        """(combined image features, final text features), both L2-normalised (model_v.py:268-338)."""
        original = ops.l2norm(self._backbone().image_features(pixel_values))
        final_image = self.visual_adapter.blend(original, self.alpha)
        if context_features is not None and context_features.nelement() > 0:
            if context_features.shape[-1] != self.text_feature_dim:
                print(f"Warning: Context feature dimension mismatch. Expected {self.text_feature_dim}, "
                      f"got {context_features.shape[-1]}. Skipping context.")
                combined = final_image
            else:
                final_ctx = self.context_adapter.blend(context_features.float().contiguous(), self.gamma)
                combined = ops.l2norm(final_image, final_ctx)  # normalise((a+b)/2) == normalise(a+b)
        else:
            combined = final_image
        if self.training or getattr(self, "adapted_emotion_embedding_tensor", None) is None:
            final_text = self.text_adapter.blend(self.emotion_embedding_tensor.clone().detach(), self.beta)
        else:
            final_text = self.adapted_emotion_embedding_tensor
        return combined, final_text

    def forward(self, pixel_values, context_features=None, use_adapters_for_training=True):
        """logits = temperature * combined f_txt^T (reference: model_v.py:260-343)."""
        combined, final_text = self.features(pixel_values, context_features)
        temperature = float(self.model.logit_scale.detach().exp().item())
        return ops.scaled_similarity(combined, final_text.contiguous(), temperature)

    def loss(self, pixel_values, labels, context_features=None):
        """nn.CrossEntropyLoss()(self(pixel_values, ctx), labels) fused into the class head (main.py:78-84)."""
        combined, final_text = self.features(pixel_values, context_features)
        temperature = float(self.model.logit_scale.detach().exp().item())
        hard = labels if labels.dtype == torch.int64 else None
        soft = labels.float().contiguous() if labels.dtype != torch.int64 else None
        loss, logits = ops.class_head_loss(combined, final_text.contiguous(), temperature, labels=hard, soft_labels=soft)
        return loss, logits

    def predict_probs(self, pixel_values, context_features=None):
        self.eval()
        with torch.no_grad():
            combined, final_text = self.features(pixel_values, context_features)
            temperature = float(self.model.logit_scale.detach().exp().item())
            probs, _ = ops.class_head_probs(combined, final_text.contiguous(), temperature)
        return probs

    def get_trainable_parameters(self):
        params = []
        params.extend(list(self.visual_adapter.parameters()))
        params.extend(list(self.text_adapter.parameters()))
        params.extend(list(self.context_adapter.parameters()))
        return params
