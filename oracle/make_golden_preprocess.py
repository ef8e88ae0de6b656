"""Writes tests/golden/resize_cv2.npz: cv2.resize (INTER_LINEAR, uint8) outputs for seeded random frames.
Run in the build container (cv2 is installed here, not assumed on the GPU box)."""
import numpy as np, cv2
from pathlib import Path
rng = np.random.default_rng(7)
cases = {}
for k, (hs, ws, hd, wd) in enumerate([(37, 53, 32, 32), (48, 64, 32, 48), (20, 28, 32, 32), (96, 128, 64, 64), (33, 33, 32, 32)]):
    src = rng.integers(0, 256, (hs, ws, 3), dtype=np.uint8)
    cases[f"src{k}"] = src
    cases[f"dst{k}"] = cv2.resize(src, (wd, hd))
out = Path(__file__).resolve().parent.parent / "tests" / "golden" / "resize_cv2.npz"
np.savez_compressed(out, **cases)
print(out, out.stat().st_size)
