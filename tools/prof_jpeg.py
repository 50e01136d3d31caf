"""One JPEG encode of 64 pages 678x960 (for ncu launch lists / captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ocr_system_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x = ops.resize_if_needed(ops.synth_pages(n, 3508, 2480, 0), 960)
enc = ops.JpegEncoder()
for _ in range(2):
    files, sizes = enc.encode(x, 95, True)
print(sizes[:4])
