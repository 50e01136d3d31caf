import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ocr_system_b200 import ops
from ocr_system_b200.pipeline import PagePipeline
pages = ops.synth_pages(64, 3508, 2480, 0)
pipe = PagePipeline(max_dimension=960)
for _ in range(3): r = pipe.run_device(pages)
ref = r.angles.copy(); refpages = r.pages.clone()
torch.cuda.synchronize()
K = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K): r = pipe.run_device(pages); del r
e1.record(); torch.cuda.synchronize()
print("sequential ms/step", e0.elapsed_time(e1) / K)
for r in pipe.run_device_stream([pages] * 3): pass
torch.cuda.synchronize()
e0.record()
last = None
for r in pipe.run_device_stream([pages] * K):
    last = (r.angles, r.pages)                # no synchronisation inside the stream, earlier results are released
    del r
e1.record(); torch.cuda.synchronize()
ok = np.array_equal(last[0], ref) and torch.equal(last[1], refpages)
print("pipelined  ms/step", e0.elapsed_time(e1) / K, "identical", ok)
