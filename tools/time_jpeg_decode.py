"""Times the device JPEG decoder on a batch of synthetic A4 pages (files made by the device encoder, which writes
Pillow's byte stream).  python tools/time_jpeg_decode.py [n_pages] [quality]"""
import io
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocr_system_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
q = int(sys.argv[2]) if len(sys.argv) > 2 else 75
torch.cuda.set_device(0)
pages = ops.synth_pages(n, 3508, 2480, seed0=0)
files = []
for i in range(0, n, 16):
    files += ops.jpeg_encode(pages[i:i + 16], quality=q)
dec = ops.JpegDecoder()
blob, offs = dec.pack(files)
blob = blob.clone().pin_memory()
out, status = dec.decode(blob, offs)
torch.cuda.synchronize()
assert int(status.abs().sum()) == 0
from PIL import Image
ref = np.asarray(Image.open(io.BytesIO(files[1])))
assert np.array_equal(out[1].cpu().numpy(), ref), "device decode differs from Pillow"
t0 = time.perf_counter(); Image.open(io.BytesIO(files[1])).load(); t_pil = time.perf_counter() - t0
reps = 20
evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
evs[0].record()
for i in range(reps):
    out, status = dec.decode(blob, offs)
    evs[i + 1].record()
torch.cuda.synchronize()
per = [evs[i].elapsed_time(evs[i + 1]) for i in range(reps)]
print("per-iteration ms:", " ".join(f"{x:.2f}" for x in per))
ms = sorted(per)[len(per) // 2]
if os.environ.get("JD_PROFILE"):
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            out, status = dec.decode(blob, offs)
        torch.cuda.synchronize()
    rows = {}
    for e in prof.events():
        if e.device_type.name == "CUDA":
            rows.setdefault(e.name[:60], []).append(e.device_time)
    for k, v in sorted(rows.items(), key=lambda kv: -sum(kv[1])):
        print(f"  {k:60s} n={len(v):3d} avg={sum(v)/len(v):9.1f} us  min={min(v):9.1f} max={max(v):9.1f}")
print(json.dumps({"pages": n, "quality": q, "file_bytes_total": int(offs[-1]), "ms_per_batch_incl_h2d_median": round(ms, 3), "ms_max": round(max(per), 3),
                  "pages_per_s": round(n / ms * 1e3, 1), "pillow_ms_per_page_one_core": round(t_pil * 1e3, 1),
                  "raster_GBps": round(n * 3508 * 2480 * 3 / ms / 1e6, 1)}))
