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
reps = 10
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    out, status = dec.decode(blob, offs)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
print(json.dumps({"pages": n, "quality": q, "file_bytes_total": int(offs[-1]), "ms_per_batch_incl_h2d": round(ms, 3),
                  "pages_per_s": round(n / ms * 1e3, 1), "pillow_ms_per_page_one_core": round(t_pil * 1e3, 1),
                  "raster_GBps": round(n * 3508 * 2480 * 3 / ms / 1e6, 1)}))
