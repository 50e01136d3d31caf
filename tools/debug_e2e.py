import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ocr_system_b200 import ops
from ocr_system_b200.pipeline import PagePipeline
dev = torch.device("cuda:0")
pages = ops.synth_pages(64, 3508, 2480, 0)
host_in = torch.empty(pages.shape, dtype=torch.uint8, pin_memory=True); host_in.copy_(pages); torch.cuda.synchronize()
pipe = PagePipeline(max_dimension=960, device=dev)
for _ in range(3): pipe.run_device(pages)
torch.cuda.synchronize()
# pure H2D
t = time.perf_counter(); x = host_in.to(dev, non_blocking=True); torch.cuda.synchronize(); print("H2D ms", (time.perf_counter()-t)*1e3)
t = time.perf_counter(); pipe.run_device(pages); torch.cuda.synchronize(); print("run_device ms", (time.perf_counter()-t)*1e3)
# H2D concurrently with run_device
cs = torch.cuda.Stream()
buf = torch.empty_like(pages)
t = time.perf_counter()
with torch.cuda.stream(cs): buf.copy_(host_in, non_blocking=True)
t1 = time.perf_counter()
pipe.run_device(pages); torch.cuda.synchronize()
print("overlapped H2D + run_device ms", (time.perf_counter()-t)*1e3, "issue ms", (t1-t)*1e3)
# stream API
t = time.perf_counter(); n = 0
for out, res, h2d, d2h in pipe.run_host_stream([host_in] * 6):
    n += 1; print("  batch", n, "t=", round((time.perf_counter()-t)*1e3, 1))
print("---- second call (persistent buffers), K=5 ----")
torch.cuda.synchronize()
t = time.perf_counter(); n = 0
for out, res, h2d, d2h in pipe.run_host_stream([host_in] * 5):
    n += 1; print("  batch", n, "t=", round((time.perf_counter()-t)*1e3, 1))
print("---- third call, K=5, with event timing like bench ----")
f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
f0.record()
for out, res, h2d, d2h in pipe.run_host_stream([host_in] * 5): pass
f1.record(); torch.cuda.synchronize(); print("event ms/step", f0.elapsed_time(f1)/5)
