"""Per-stage device timings of the page chain (CUDA events on the launch stream)."""
import argparse
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ocr_system_b200 import ops


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        r = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return r, float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--md", type=int, default=960)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    n, md = a.n, a.md
    H, W = 3508, 2480
    pages = ops.synth_pages(n, H, W, 0)
    torch.cuda.synchronize()
    tw, th = ops.target_size(W, H, md)
    res = {}

    def rec(name, fn, nbytes):
        r, med, mn = timed(fn, a.iters)
        res[name] = {"ms": med, "ms_min": mn, "GBps": nbytes / med / 1e6}
        print(f"{name:28s} {med:9.3f} ms (min {mn:8.3f})  {nbytes/med/1e6:9.1f} GB/s", flush=True)
        return r

    small = rec("resize_lanczos", lambda: ops.resize_lanczos(pages, tw, th), n * (H * W * 3 + th * tw * 3))
    rec("gray_pil_fullres", lambda: ops.gray_pil(pages), n * H * W * 4)
    px = n * th * tw
    rec("gray_cv", lambda: ops.gray_cv(small), px * 4)
    edges = rec("canny", lambda: ops.canny(small), px * 4)
    rec("ppht", lambda: ops.hough_lines_p(edges), px)
    rec("deskew_total", lambda: ops.deskew(small), px * 8)
    mats = np.stack([ops.rotation_matrix(tw // 2, th // 2, 1.0).reshape(6)] * n)
    rec("warp_affine", lambda: ops.warp_affine_cubic(small, mats), px * 6)
    rec("contrast_mean", lambda: ops.contrast_mean(small), px * 3)
    rec("contrast_sharpness_fused", lambda: ops.contrast_sharpness(small, 1.2, 1.1), px * 9)
    rec("sharpness", lambda: ops.enhance_sharpness(small, 1.1), px * 6)
    rec("median3", lambda: ops.median3(small), px * 6)
    rec("adaptive_binarize", lambda: ops.adaptive_binarize(small), px * 4)
    rec("det_resize_normalize", lambda: ops.det_resize_normalize(small), px * 3 + n * 3 * 960 * 672 * 4)
    lines, nl = ops.hough_lines_p(edges)
    print("edge px/page", (edges != 0).sum().item() / n, "lines/page", nl.float().mean().item())
    json.dump(res, open(f"gpurun_out/stages_md{md}_n{n}.json", "w"), indent=1)


if __name__ == "__main__":
    os.makedirs("gpurun_out", exist_ok=True)
    main()
