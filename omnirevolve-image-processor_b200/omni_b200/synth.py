"""Seeded synthetic inputs of SURVEY.md 8d / BASELINE.md 3 (host side; needs cv2 for INTER_CUBIC)."""
import cv2
import numpy as np


def synth(H: int, W: int, seed: int, cell: int = 32) -> np.ndarray:
    """Smooth random colour regions + +-12 noise, HxWx3 u8 (treated as BGR)."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (max(1, H // cell), max(1, W // cell), 3), np.uint8)
    img = cv2.resize(base, (W, H), interpolation=cv2.INTER_CUBIC)
    noise = rng.integers(-12, 13, img.shape, dtype=np.int16)
    return np.clip(img.astype(np.int16) + noise, 0, 255).astype(np.uint8)


def layer_names(K: int):
    return [f"layer_{i:02d}" for i in range(K)]


def layer_colors(K: int):
    return [[(37 * i) % 256, (91 * i) % 256, (53 * i + 40) % 256] for i in range(K)]
