"""Reading-order row: 256 pages x ~500 boxes (the DBPostProcess config's box count).
GPU kernel time with device-resident inputs (CUDA events), the host-list API end to end, and the CPU
restatement (oracle/reading_order.py = the reference's pure-Python algorithm) on a sample."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import numpy as np, torch
from reading_pages import page
from ocr_system_b200 import ops, ocr_postprocessor as P
from oracle import reading_order as R

NP, NB = 256, 500
pages = [page(1000 + i, NB, "grid") for i in range(NP)]
off = np.zeros(NP + 1, np.int32); np.cumsum([len(p) for p in pages], out=off[1:])
boxes = torch.from_numpy(np.array([it[0] for p in pages for it in p], np.float64)).cuda()
conf = torch.from_numpy(np.array([it[2] for p in pages for it in p], np.float64)).cuda()
offt = torch.from_numpy(off)
for _ in range(3):
    ops.reading_order(boxes, conf, offt, 0.5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 20
for _ in range(K):
    out = ops.reading_order(boxes, conf, offt, 0.5)
e1.record(); torch.cuda.synchronize()
gpu_ms = e0.elapsed_time(e1) / K
t = time.perf_counter(); merged = P.process_ocr_results_batch(pages, 0.5); api_ms = (time.perf_counter() - t) * 1e3
t = time.perf_counter()
S = 16
for p in pages[:S]:
    R.reading_order([it[0] for it in p], [it[2] for it in p], 0.5)
cpu_ms = (time.perf_counter() - t) * 1e3 / S
res = {"pages": NP, "boxes_per_page": NB, "gpu_ms_per_batch": round(gpu_ms, 4), "gpu_pages_per_s": round(NP / gpu_ms * 1e3),
       "host_api_ms_per_batch": round(api_ms, 2), "cpu_python_ms_per_page": round(cpu_ms, 3),
       "cpu_pages_per_s_1core": round(1e3 / cpu_ms), "lines_page0": len(merged[0])}
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/reading_order_timing.json", "w"))
