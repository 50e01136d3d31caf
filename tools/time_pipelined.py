"""Sequential vs stream-pipelined device steps (PagePipeline.run_device / run_device_stream), several K."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ocr_system_b200 import ops
from ocr_system_b200.pipeline import PagePipeline
pages = ops.synth_pages(64, 3508, 2480, 0)
pipe = PagePipeline(max_dimension=960)
for _ in range(3): r = pipe.run_device(pages)
ref = r.angles.copy(); refpages = r.pages.clone(); del r
for r in pipe.run_device_stream([pages] * 3): del r
torch.cuda.synchronize()
for K in (5, 10, 20, 40, 20, 10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K): r = pipe.run_device(pages); del r
    e1.record(); torch.cuda.synchronize()
    seq = e0.elapsed_time(e1) / K
    e0.record()
    last = None
    for r in pipe.run_device_stream([pages] * K):
        last = (r.angles, r.pages); del r
    e1.record(); torch.cuda.synchronize()
    ok = np.array_equal(last[0], ref) and torch.equal(last[1], refpages)
    print(f"K={K:3d} sequential {seq:6.2f} ms/step   pipelined {e0.elapsed_time(e1) / K:6.2f} ms/step  identical {ok}  mem {torch.cuda.memory_reserved() / 2**30:.1f} GiB")
