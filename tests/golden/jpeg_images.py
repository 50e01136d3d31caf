"""Seeded small RGB test images shared by the JPEG golden generator and the tests."""
import numpy as np

SHAPES = [(16, 16), (50, 70), (93, 127), (8, 200), (121, 33)]
KINDS = ["smooth", "noise", "text", "photo"]


def image(kind: str, h: int, w: int, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed * 1000 + h * 7 + w)
    yy, xx = np.mgrid[0:h, 0:w]
    if kind == "smooth":
        a = np.stack([(xx * 3 + yy) % 256, (yy * 2) % 256, (xx + yy * 5) % 256], -1)
    elif kind == "noise":
        a = rng.integers(0, 256, (h, w, 3))
    elif kind == "text":
        a = np.full((h, w, 3), 255)
        a[(yy // 3 + xx // 5) % 4 == 0] = 20
    else:  # photo-like: low-frequency colour field plus mild noise
        base = 128 + 90 * np.sin(xx / 9.0 + seed) * np.cos(yy / 7.0)
        a = np.stack([base, base * 0.8 + 20, 255 - base], -1) + rng.normal(0, 6, (h, w, 3))
        a = np.clip(np.rint(a), 0, 255)
    return np.ascontiguousarray(a.astype(np.uint8))
