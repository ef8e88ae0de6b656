"""TEST INFRASTRUCTURE (CPU oracle) for the on-GPU frame preprocessing + patch extraction.

Restates what the reference does to a decoded video frame before a model could see it
(process_video.py:14-29): `cv2.cvtColor(BGR2RGB)` -> `cv2.resize(frame, size)` (INTER_LINEAR on uint8, i.e. OpenCV's
11-bit fixed-point bilinear) -> `ToTensor()` (/255) -> `Normalize(mean, std)` -> stack to [C, T, H, W];
and the patch extraction of HF CLIPVisionEmbeddings (HF:209, Conv2d with kernel = stride = patch as im2col).
Pinned against cv2 itself (tests/golden/resize_cv2.npz, written by oracle/make_golden_preprocess.py in the build
container where cv2 is installed).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this.
"""
from __future__ import annotations

import numpy as np

IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)          # process_video.py:24
CLIP_MEAN, CLIP_STD = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)  # CLIPImageProcessor

COEF_BITS = 11
COEF_ONE = 1 << COEF_BITS


def linear_coeffs(src: int, dst: int):
    """OpenCV resize.cpp, INTER_LINEAR: source index and the two 11-bit weights of every destination index."""
    scale = src / dst
    idx = np.empty(dst, np.int64)
    w = np.empty((dst, 2), np.int64)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)  # resize.cpp: fx = (float)((dx + 0.5) * scale_x - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if s < 0:
            s, f = 0, np.float32(0.0)
        if s >= src - 1:
            s, f = src - 1, np.float32(0.0)
        idx[d] = s
        # saturate_cast<short>(x * 2048) rounds half to even (cvRound); arithmetic in float32 as in OpenCV
        w[d, 0] = int(np.rint(np.float32(np.float32(1.0) - f) * np.float32(COEF_ONE)))
        w[d, 1] = int(np.rint(f * np.float32(COEF_ONE)))
    return idx, w


def resize_bilinear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """cv2.resize(img, (out_w, out_h)) for uint8 HWC input (INTER_LINEAR), bit for bit."""
    assert img.dtype == np.uint8 and img.ndim == 3
    h, w, _ = img.shape
    if (h, w) == (out_h, out_w):
        return img.copy()
    xi, xw = linear_coeffs(w, out_w)
    yi, yw = linear_coeffs(h, out_h)
    src = img.astype(np.int64)
    x1 = np.minimum(xi + 1, w - 1)
    # horizontal pass: S = src[sx] * a0 + src[sx + 1] * a1
    hor = src[:, xi, :] * xw[None, :, 0, None] + src[:, x1, :] * xw[None, :, 1, None]  # [h, out_w, c]
    y1 = np.minimum(yi + 1, h - 1)
    s0, s1 = hor[yi], hor[y1]
    b0, b1 = yw[:, 0][:, None, None], yw[:, 1][:, None, None]
    out = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def normalise(frames_u8: np.ndarray, mean, std) -> np.ndarray:
    """ToTensor + Normalize: uint8 [..., H, W, 3] -> float32 [..., 3, H, W]."""
    x = frames_u8.astype(np.float32) / np.float32(255.0)
    x = (x - np.asarray(mean, np.float32)) / np.asarray(std, np.float32)
    return np.moveaxis(x, -1, -3)


def preprocess_frames(frames_u8: np.ndarray, out_h: int, out_w: int, mean, std, bgr: bool = False) -> np.ndarray:
    """[N, Hs, Ws, 3] uint8 -> [N, 3, out_h, out_w] float32 (BGR2RGB optional, resize, /255, normalise)."""
    outs = []
    for f in frames_u8:
        if bgr:
            f = f[..., ::-1]
        outs.append(resize_bilinear_u8(np.ascontiguousarray(f), out_h, out_w))
    return normalise(np.stack(outs), mean, std)


def patches(pixels: np.ndarray, patch: int) -> np.ndarray:
    """[N, 3, H, W] -> [N * (H/p) * (W/p), 3*p*p] in the column order of patch_embedding.weight.view(D, -1)."""
    n, c, h, w = pixels.shape
    gh, gw = h // patch, w // patch
    x = pixels.reshape(n, c, gh, patch, gw, patch).transpose(0, 2, 4, 1, 3, 5)
    return x.reshape(n * gh * gw, c * patch * patch)


def video_mean_pool(frame_features: np.ndarray, clips: int, frames: int) -> np.ndarray:
    """get_image_features(frames).view(B, T, P).mean(1) (before normalisation)."""
    return frame_features.reshape(clips, frames, -1).mean(1)
