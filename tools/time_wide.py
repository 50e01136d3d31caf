"""Device timings of the wide (non-Hough) kernels of the page chain at the bench geometry (64 pages 678x960),
CUDA events on the launch stream, median of --iters launches after warm-up.  Writes gpurun_out/wide_kernels.json."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ocr_system_b200 import ops


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    n = a.n
    H, W = 3508, 2480
    pages = ops.synth_pages(n, H, W, 0)
    tw, th = ops.target_size(W, H, 960)
    small = ops.resize_lanczos(pages, tw, th)
    del pages
    gray = ops.gray_pil(small)
    px = n * th * tw
    res = {}

    def rec(name, fn, nbytes):
        med, mn = timed(fn, a.iters)
        res[name] = {"ms": med, "ms_min": mn, "GBps": nbytes / med / 1e6}
        print(f"{name:28s} {med:9.4f} ms (min {mn:8.4f})  {nbytes / med / 1e6:9.1f} GB/s", flush=True)

    mats = np.stack([ops.rotation_matrix(tw // 2, th // 2, 1.0 + 0.01 * i).reshape(6) for i in range(n)])
    rec("warp_affine_rgb", lambda: ops.warp_affine_cubic(small, mats), px * 6)
    rec("canny", lambda: ops.canny(small, 50, 150), px * 4)
    rec("gray_pil", lambda: ops.gray_pil(small), px * 4)
    rec("adaptive_binarize_gray", lambda: ops.adaptive_binarize(gray, 2), px * 2)
    rec("det_resize_normalize", lambda: ops.det_resize_normalize(small, 960), px * 3 + n * 3 * 960 * 672 * 4)
    rec("contrast_sharpness_fused", lambda: ops.contrast_sharpness(small, 1.2, 1.1), px * 9)
    rec("median3", lambda: ops.median3(small), px * 6)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/wide_kernels{a.tag}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
