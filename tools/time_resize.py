"""Resize kernel timing (64 A4 pages -> max side 960), CUDA events, median of 20; set LUMINA_RESIZE_STAGED=1 /
LUMINA_RESIZE_DP4A=1 for the older kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ocr_system_b200 import ops
pages = ops.synth_pages(64, 3508, 2480, 0)
for target in (960, 2000):
    for _ in range(3):
        small = ops.resize_if_needed(pages, target)
    ts = []
    for _ in range(20):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); small = ops.resize_if_needed(pages, target); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    gb = (pages.numel() + small.numel()) / 1e9
    print(f"target {target}: median {ts[10]:.3f} ms  min {ts[0]:.3f} ms  {gb / ts[10] * 1e3:.0f} GB/s  out {tuple(small.shape)}")
