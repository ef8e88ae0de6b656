"""B200-native mirror of the reference's Track-T model (reference: model_t.py).

CLIP-Adapter style fine-tuning on the pooled, L2-normalised CLIP embeddings: a ReLU bottleneck per branch,
alpha/beta residual blend, re-normalisation, `temperature * I T^T`, cross-entropy against class prompts
(model_t.py:157-187), plus the predict paths (model_t.py:213-298) and the zero-shot baseline (:300-404).

Kernel mapping
  frozen image tower + post_layernorm + projection   NativeClipTowers.image_features   (model_t.py:158-160)
  adapter -> blend -> renormalise                     ONE fused kernel, post = BLEND_L2  (model_t.py:163-181)
  logits + cross-entropy + gradients                  vlmclip_class_head                 (model_t.py:184-187)
  Adam                                                FusedAdamW(weight_decay=0, no clip) (model_t.py:141-145,190-192)
  predict / predict_with_all_descriptions             class head forward with softmax, group max over 5 prompts
"""
from __future__ import annotations

import torch
from torch import nn

from . import _native as N
from . import ops
from .constants import EMOTIONS, get_emotion_descriptions
from .model_m import _load_clip, _load_processor
from .towers import NativeClipTowers

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")


class _ReluBottleneck(nn.Module):
    def __init__(self, input_dim, bottleneck_dim):
        super().__init__()
        self.fc1 = nn.Linear(input_dim, bottleneck_dim)
        self.fc2 = nn.Linear(bottleneck_dim, input_dim)
        self.relu = nn.ReLU()

    def forward(self, x):
        """fc2(relu(fc1(x))) (model_t.py:21-22)."""
        x2 = x.reshape(-1, x.shape[-1]).float().contiguous()
        y = ops.adapter(x2, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, act=N.ACT_RELU,
                        post=N.POST_PLAIN)
        return y.view(*x.shape)

    def blend(self, x, ratio: float, hmask=None):
        """normalise(ratio * self(x) + (1 - ratio) * x) in one kernel (model_t.py:163-169)."""
        x2 = x.reshape(-1, x.shape[-1]).float().contiguous()
        return ops.adapter(x2, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, act=N.ACT_RELU,
                           post=N.POST_BLEND_L2, alpha=float(ratio), hmask=hmask)


class VisualAdapter(_ReluBottleneck):
    """Adapter module for the visual branch of CLIP (reference: model_t.py:13-22)."""


class TextAdapter(_ReluBottleneck):
    """Adapter module for the text branch of CLIP (reference: model_t.py:24-33)."""


class _ClipHolder:
    """Shared by CLIPAdapter and ZeroShotEmotionRecognition: frozen CLIP weights + native towers + prompt encoding."""

    def _init_clip(self, model_name, clip, processor):
        self.model = (clip if clip is not None else _load_clip(model_name)).to(device)
        self.processor = processor if processor is not None else (None if clip is not None else _load_processor(model_name))
        for param in self.model.parameters():
            param.requires_grad = False
        self._towers = None

    def _backbone(self) -> NativeClipTowers:
        if self._towers is None:
            self._towers = NativeClipTowers(self.model, next(self.model.parameters()).device)
        return self._towers

    def _image_features_normalised(self, pixel_values):
        feats = self._backbone().image_features(pixel_values.to(next(self.model.parameters()).device))
        return ops.l2norm(feats)

    def _encode_prompts(self, store_mean_as: str):
        """Per-prompt text embeddings (batch 1, variable length — model_t.py:84-109) and their per-class means."""
        if self.processor is None:
            raise N.NativeError("no CLIPProcessor available offline: pass processor=..., or set the class embeddings "
                                "directly (emotion_embedding_tensor / emotion_text_features_per_description)")
        print("Encoding emotion descriptions...")
        per_class_mean = {}
        self.emotion_text_features_per_description = {}
        dev = next(self.model.parameters()).device
        for emotion, descriptions in self.emotion_descriptions.items():
            feats = []
            for description in descriptions:
                enc = self.processor(text=[description], padding=True, truncation=True, return_tensors="pt")
                f = self._backbone().text_features(enc["input_ids"].to(dev), enc["attention_mask"].to(dev))
                feats.append(ops.l2norm(f))
            self.emotion_text_features_per_description[emotion] = feats
            per_class_mean[emotion] = torch.cat(feats, dim=0).mean(dim=0, keepdim=True)
        setattr(self, store_mean_as, per_class_mean)
        self.emotion_embedding_tensor = torch.cat(list(per_class_mean.values()), dim=0)


class CLIPAdapter(_ClipHolder):
    """CLIP-Adapter: fine-tuning CLIP with bottleneck adapters for few-shot learning (reference: model_t.py:35-298)."""

    def __init__(self, model_name, alpha=0.2, beta=0.2, bottleneck_dim=64, *, clip=None, processor=None,
                 emotion_descriptions=None, encode=True):
        self._init_clip(model_name, clip, processor)
        self.image_feature_dim = self.model.config.projection_dim
        self.text_feature_dim = self.image_feature_dim
        self.visual_adapter = VisualAdapter(self.image_feature_dim, bottleneck_dim).to(device)
        self.text_adapter = TextAdapter(self.text_feature_dim, bottleneck_dim).to(device)
        self.alpha = alpha
        self.beta = beta
        self.emotion_descriptions = emotion_descriptions if emotion_descriptions is not None else get_emotion_descriptions()
        if encode:
            self.encode_emotion_descriptions()

    def encode_emotion_descriptions(self):
        self._encode_prompts("original_emotion_text_features")

    def update_emotion_embeddings(self):
        """Adapted class embeddings for inference (reference: model_t.py:111-129)."""
        with torch.no_grad():
            self.adapted_emotion_embedding_tensor = self.text_adapter.blend(self.emotion_embedding_tensor, self.beta)

    def train_step(self, pixel_values, labels, optimizer, temperature: float):
        """One step of model_t.py:153-192; `labels` int64 [B] or fp32 [B, C] class probabilities (soft labels)."""
        original = self._image_features_normalised(pixel_values)
        final_image = self.visual_adapter.blend(original, self.alpha)
        final_text = self.text_adapter.blend(self.emotion_embedding_tensor.clone().detach(), self.beta)
        hard = labels if labels.dtype == torch.int64 else None
        soft = labels.float().contiguous() if labels.dtype != torch.int64 else None
        loss, _ = ops.class_head_loss(final_image, final_text, temperature, labels=hard, soft_labels=soft)
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        return loss.detach()

    def train(self, train_loader, num_epochs=50, learning_rate=3e-4):
        from tqdm import tqdm

        self.visual_adapter.train()
        self.text_adapter.train()
        optimizer = ops.FusedAdamW(list(self.visual_adapter.parameters()) + list(self.text_adapter.parameters()),
                                   lr=learning_rate, weight_decay=0.0, max_grad_norm=0.0)  # optim.Adam (model_t.py:141-145)
        temperature = self.model.logit_scale.exp().item()
        for epoch in range(num_epochs):
            total = None
            batch_count = 0
            progress_bar = tqdm(train_loader, desc=f"Epoch {epoch+1}/{num_epochs}")
            for pixel_values, labels, _ in progress_bar:
                loss = self.train_step(pixel_values.to(device), labels.to(device), optimizer, temperature)
                total = loss.clone() if total is None else total + loss
                batch_count += 1
            avg_loss = (total.item() if total is not None else 0.0) / max(1, batch_count)
            print(f"Epoch {epoch+1}/{num_epochs}, Loss: {avg_loss:.4f}")
            self.update_emotion_embeddings()
        self.update_emotion_embeddings()
        self.visual_adapter.eval()
        self.text_adapter.eval()

    def _final_image_features(self, pixel_values):
        original = self._image_features_normalised(pixel_values)
        if hasattr(self, "visual_adapter"):
            return self.visual_adapter.blend(original, self.alpha)
        return original

    def predict(self, pixel_values):
        """softmax(100 * f_img f_txt^T) (reference: model_t.py:213-242)."""
        with torch.no_grad():
            f_img = self._final_image_features(pixel_values)
            emb = getattr(self, "adapted_emotion_embedding_tensor", None)
            if emb is None:
                emb = self.emotion_embedding_tensor
            probs, _ = ops.class_head_probs(f_img, emb.float().contiguous(), 100.0)
        return probs

    def predict_with_all_descriptions(self, pixel_values):
        """Per-class max over the individual prompts, then softmax (reference: model_t.py:244-298) — one [B, C*G]
        similarity with a segmented max instead of C*G Python-loop matvecs."""
        with torch.no_grad():
            f_img = self._final_image_features(pixel_values)
            classes = list(self.emotion_text_features_per_description.keys())
            if list(EMOTIONS) == classes or all(e in self.emotion_text_features_per_description for e in EMOTIONS):
                classes = [e for e in EMOTIONS if e in self.emotion_text_features_per_description] or classes
            groups = {len(self.emotion_text_features_per_description[c]) for c in classes}
            if len(groups) != 1:
                raise ValueError("predict_with_all_descriptions needs the same number of prompts for every class")
            G = groups.pop()
            dev = f_img.device
            per_prompt = torch.cat([torch.cat([f.to(dev) for f in self.emotion_text_features_per_description[c]], 0)
                                    for c in classes], 0).float().contiguous()
            adapted = self.text_adapter.blend(per_prompt, self.beta)
            probs, _ = ops.class_head_probs(f_img, adapted, 100.0, group=G)
        return probs


class ZeroShotEmotionRecognition(_ClipHolder):
    """Zero-shot emotion recognition with detailed descriptions (reference: model_t.py:300-404)."""

    def __init__(self, model_name, *, clip=None, processor=None, emotion_descriptions=None, encode=True):
        self._init_clip(model_name, clip, processor)
        self.emotion_descriptions = emotion_descriptions if emotion_descriptions is not None else get_emotion_descriptions()
        if encode:
            self.encode_emotion_descriptions()

    def encode_emotion_descriptions(self):
        self._encode_prompts("emotion_text_features")

    def predict(self, pixel_values):
        with torch.no_grad():
            f_img = self._image_features_normalised(pixel_values)
            probs, _ = ops.class_head_probs(f_img, self.emotion_embedding_tensor.to(f_img.device).float().contiguous(), 100.0)
        return probs

    def predict_with_all_descriptions(self, pixel_values):
        with torch.no_grad():
            f_img = self._image_features_normalised(pixel_values)
            classes = list(self.emotion_text_features_per_description.keys())
            G = len(self.emotion_text_features_per_description[classes[0]])
            per_prompt = torch.cat([torch.cat([f.to(f_img.device) for f in self.emotion_text_features_per_description[c]], 0)
                                    for c in classes], 0).float().contiguous()
            probs, _ = ops.class_head_probs(f_img, per_prompt, 100.0, group=G)
        return probs
