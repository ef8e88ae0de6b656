"""CPU: the preprocessing oracle against cv2's own output (golden written by oracle/make_golden_preprocess.py in the
build container) and against torch's unfold; the product's coefficient tables against the oracle's."""
from pathlib import Path

import numpy as np
import torch

from oracle import preprocess_oracle as P

GOLD = Path(__file__).resolve().parent / "golden" / "resize_cv2.npz"


def test_resize_matches_cv2_golden():
    g = np.load(GOLD)
    for k in range(5):
        src, dst = g[f"src{k}"], g[f"dst{k}"]
        out = P.resize_bilinear_u8(src, dst.shape[0], dst.shape[1])
        diff = np.abs(out.astype(int) - dst.astype(int))
        down = src.shape[0] >= dst.shape[0] and src.shape[1] >= dst.shape[1]
        # bit exact when shrinking (what video frames -> 224 x 224 does); cv2's up-scaling path differs by 1 LSB on ~1 %
        assert diff.max() <= (0 if down else 1), (k, diff.max())
        assert (diff > 0).mean() < 0.02


def test_patches_match_unfold():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 3, 32, 48)).astype(np.float32)
    ref = torch.nn.functional.unfold(torch.from_numpy(x), kernel_size=16, stride=16).transpose(1, 2).reshape(-1, 3 * 256)
    assert np.array_equal(P.patches(x, 16), ref.numpy())


def test_normalise_is_totensor_normalize():
    rng = np.random.default_rng(1)
    f = rng.integers(0, 256, (3, 8, 8, 3), dtype=np.uint8)
    y = P.normalise(f, P.IMAGENET_MEAN, P.IMAGENET_STD)
    t = torch.from_numpy(f).permute(0, 3, 1, 2).float().div(255.0)
    m = torch.tensor(P.IMAGENET_MEAN).view(1, 3, 1, 1)
    s = torch.tensor(P.IMAGENET_STD).view(1, 3, 1, 1)
    assert np.array_equal(y, ((t - m) / s).numpy())


def test_product_coefficient_tables_equal_the_oracle():
    from vlm_clip_b200 import ops

    for src, dst in [(480, 224), (640, 224), (100, 224), (224, 224), (37, 32), (500, 224)]:
        idx, w = P.linear_coeffs(src, dst)
        t = ops._linear_coeffs(src, dst)
        assert (t[:, 0] == idx).all() and (t[:, 1:] == w).all()
        assert ((t[:, 1] + t[:, 2]) == 2048).all()


def test_video_mean_pool_definition():
    x = np.arange(2 * 3 * 4, dtype=np.float32).reshape(6, 4)
    assert np.allclose(P.video_mean_pool(x, 2, 3), torch.from_numpy(x).view(2, 3, 4).mean(1).numpy())
