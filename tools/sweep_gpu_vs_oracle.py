"""Randomised GPU-vs-oracle sweep over geometries the parity tests do not enumerate: random page shapes, batch
sizes and max_dimensions through every preprocessing op of the C-ABI, compared byte for byte with the CPU oracle
(test infrastructure: this tool is a checker, like tests/).  Mismatches are printed with the seed.

    python tools/sweep_gpu_vs_oracle.py --seeds 0 60
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import oracle as O
from ocr_system_b200 import ops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, nargs=2, default=[0, 40])
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    bad, checked = [], 0
    t0 = time.time()

    def t(x):
        return torch.from_numpy(np.ascontiguousarray(x)).to(dev)

    def cmp(seed, name, got, want, info):
        nonlocal checked
        checked += 1
        g = got.cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
        if g.shape != np.asarray(want).shape or not np.array_equal(g, want):
            bad.append((seed, name, info))
            print("MISMATCH", seed, name, info, flush=True)

    for seed in range(*a.seeds):
        rng = np.random.default_rng(seed)
        h, w = int(rng.integers(40, 1500)), int(rng.integers(40, 1500))
        if seed % 7 == 0:
            h, w = int(rng.integers(1500, 3600)), int(rng.integers(1500, 2600))
        md = int(rng.integers(32, max(h, w)))
        n = int(rng.integers(1, 4))
        info = dict(h=h, w=w, md=md, n=n)
        kind = seed % 3
        imgs = [O.synth_page(h, w, seed * 10 + i) if kind else rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for i in range(n)]
        x = t(np.stack(imgs))
        tw, th = O.target_size(w, h, md)
        small = ops.resize_if_needed(x, md)
        ref_small = [im if (tw, th) == (w, h) else O.resize_lanczos(im, tw, th) for im in imgs]
        cmp(seed, "resize_rgb", small, np.stack(ref_small), info)
        g_full = [O.gray_pil(im) for im in imgs]
        cmp(seed, "gray_pil", ops.gray_pil(x), np.stack(g_full), info)
        cmp(seed, "resize_L", ops.resize_if_needed(t(np.stack(g_full)), md),
            np.stack([g if (tw, th) == (w, h) else O.resize_lanczos(g, tw, th) for g in g_full]), info)
        rs = np.stack(ref_small)
        xs = t(rs)
        cmp(seed, "contrast+sharpness", ops.contrast_sharpness(xs, 1.2, 1.1), np.stack([O.sharpness(O.contrast(im, 1.2), 1.1) for im in rs]), info)
        cmp(seed, "median3", ops.median3(xs), np.stack([O.median3(im) for im in rs]), info)
        gs = np.stack([O.gray_pil(im) for im in rs])
        for mode in ("plain", "avx2"):
            ref_b = np.stack([O.adaptive_gauss11(g, 2, cv_dispatch=mode) for g in gs])
            cmp(seed, "adaptive_" + mode, ops.adaptive_binarize(t(gs), 2, cv_dispatch=mode), ref_b, info)
            cmp(seed, "adaptive_rgb_" + mode, ops.adaptive_binarize(xs, 2, cv_dispatch=mode), ref_b, info)
        if th >= 8 and tw >= 8:
            edges = ops.canny(xs, 50, 150)
            ref_e = np.stack([O.canny(O.gray_cv(im), 50, 150) for im in rs])
            cmp(seed, "canny", edges, ref_e, info)
            out, angles = ops.deskew(xs)
            for i in range(n):
                rimg, rang, rlines = O.deskew(rs[i])
                cmp(seed, "deskew_raster", out[i], rimg, info)
                checked += 1
                if angles[i] != rang:
                    bad.append((seed, "angle", info))
                    print("MISMATCH", seed, "angle", repr(angles[i]), repr(rang), info, flush=True)
            det, sl = ops.det_resize_normalize(xs, 960)
            for i in range(n):
                rd, rsl = O.det_resize_normalize(rs[i], 960)
                cmp(seed, "det_normalize", det[i], rd, info)
    print(json.dumps({"seeds": a.seeds, "checked": checked, "mismatches": len(bad), "seconds": round(time.time() - t0, 1)}))
    with open("gpurun_out/sweep_gpu_vs_oracle.json", "w") as f:
        json.dump({"seeds": a.seeds, "checked": checked, "mismatches": bad}, f)


if __name__ == "__main__":
    os.makedirs("gpurun_out", exist_ok=True)
    main()
