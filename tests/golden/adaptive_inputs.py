"""Seeded inputs shared by make_adaptive_golden.py and the tests: gray planes on which OpenCV's optimised and plain
adaptiveThreshold paths are known to differ (selected by the generator), of every width residue mod 8."""
import numpy as np


def plane(seed: int):
    """kind 0: nearly flat noise (pixel - mean hovers around the threshold), narrow and tall: many remainder columns;
    kind 1: uniform noise; kind 2: synthetic text page (needs the oracle's generator)."""
    rng = np.random.default_rng(1000 + seed)
    kind = seed % 3
    if kind == 0:
        w, h = int(rng.integers(11, 48)), 4000
        base = int(rng.integers(20, 230))
        return (base + rng.integers(-3, 4, (h, w))).astype(np.uint8)
    h, w = int(rng.integers(300, 1100)), int(rng.integers(300, 1100))
    if kind == 1:
        return rng.integers(0, 256, (h, w), dtype=np.uint8)
    import oracle as O

    return O.gray_pil(O.synth_page(h, w, seed))
