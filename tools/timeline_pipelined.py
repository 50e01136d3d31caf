"""Timeline (CUDA events) of the stages of consecutive batches in PagePipeline.run_device_stream."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ocr_system_b200 import ops
from ocr_system_b200.pipeline import PagePipeline
pages = ops.synth_pages(64, 3508, 2480, 0)
pipe = PagePipeline(max_dimension=960)
for r in pipe.run_device_stream([pages] * 3): pass
torch.cuda.synchronize()
base = torch.cuda.Event(enable_timing=True); base.record(); 
timers = []
for r in pipe.run_device_stream([pages] * 6, profile=True):
    timers.append(r.timer)
torch.cuda.synchronize()
for i, t in enumerate(timers):
    print(f"batch {i}: " + "  ".join(f"{n}[{base.elapsed_time(a):6.1f}-{base.elapsed_time(b):6.1f}]" for n, a, b in t.marks))
