"""One HoughLinesP call on 8 resized synthetic pages (for ncu captures of the PPHT kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ocr_system_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
pages = ops.synth_pages(n, 3508, 2480, 0)
small = ops.resize_if_needed(pages, 960)
edges = ops.canny(small)
for _ in range(2):
    lines, nl = ops.hough_lines_p(edges)
torch.cuda.synchronize()
print(nl.tolist())
