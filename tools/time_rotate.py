"""Host-side timing of the deskew decision step (PagePipeline._rotate) piece by piece."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ocr_system_b200 import ops
n = 64
pages = ops.synth_pages(n, 3508, 2480, 0)
x = ops.resize_if_needed(pages, 960)
edges = ops.canny(x)
lines, nlines = ops.hough_lines_p(edges)
torch.cuda.synchronize()
for it in range(3):
    t = [time.perf_counter()]
    nl = nlines.cpu().numpy(); t.append(time.perf_counter())
    keep = int(nl.max(initial=0)); lh = lines[:, :max(keep, 1)].cpu().numpy(); t.append(time.perf_counter())
    angles, mats, apply = ops.deskew_decide(lh, nl, x.shape[1], x.shape[2]); t.append(time.perf_counter())
    y = ops.warp_affine_cubic(x, mats, apply); t.append(time.perf_counter())
    torch.cuda.synchronize(); t.append(time.perf_counter())
    print("nl.cpu %.3f  lines.cpu %.3f  decide %.3f  warp launch %.3f  warp sync %.3f ms" % tuple((b - a) * 1e3 for a, b in zip(t, t[1:])), "keep", keep)
