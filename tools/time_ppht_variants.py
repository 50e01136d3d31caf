import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ocr_system_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
pages = ops.synth_pages(n, 3508, 2480, 0)
small = ops.resize_if_needed(pages, 960)
edges = ops.canny(small)
print("edge px per page min/mean/max", (edges != 0).flatten(1).sum(1).min().item(), (edges != 0).flatten(1).sum(1).float().mean().item(), (edges != 0).flatten(1).sum(1).max().item())
for var in ("", "cluster", "l2"):
    if var: os.environ["LUMINA_PPHT"] = var
    else: os.environ.pop("LUMINA_PPHT", None)
    for m in (1, 16, 32, 48, 64):
        if m > n: continue
        e = edges[:m].contiguous()
        ts = []
        for it in range(3):
            torch.cuda.synchronize(); t = time.time(); lines, nl = ops.hough_lines_p(e); torch.cuda.synchronize(); ts.append(time.time() - t)
        print(f"variant={var or 'lm':8s} pages={m:3d} ms={min(ts)*1e3:8.2f}  lines/page={nl.float().mean().item():.0f}", flush=True)
