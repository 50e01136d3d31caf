"""How fast can a PIL RGB page reach a pinned staging buffer?  np.asarray (Pillow's raw encoder, holds the GIL)
vs the Arrow zero-copy view of Pillow's RGBX storage (needs PILLOW block size >= image size) + threaded copy."""
import os, sys, time
import numpy as np, torch, pyarrow as pa
from concurrent.futures import ThreadPoolExecutor
from PIL import Image
Image.core.set_block_size(64 * 1024 * 1024)
n, h, w = 32, 3508, 2480
a = np.random.default_rng(0).integers(0, 256, (h, w, 3), dtype=np.uint8)
ims = [Image.fromarray(a).copy() for _ in range(n)]
st3 = torch.empty((n, h, w, 3), dtype=torch.uint8, pin_memory=True).numpy()
st4 = torch.empty((n, h, w, 4), dtype=torch.uint8, pin_memory=True).numpy()
def via_asarray(i): st3[i] = np.asarray(ims[i])
def via_arrow(i): np.copyto(st4[i], pa.array(ims[i]).flatten().to_numpy(zero_copy_only=True).reshape(h, w, 4))
for name, fn in (("np.asarray", via_asarray), ("arrow RGBX", via_arrow)):
    fn(0)
    for nt in (1, 4, 8, 16):
        with ThreadPoolExecutor(nt) as ex:
            t = time.perf_counter(); list(ex.map(fn, range(n))); dt = time.perf_counter() - t
        print(f"{name:12s} threads={nt:2d}: {dt / n * 1e3:6.2f} ms/page")
print("equal", np.array_equal(st4[3][..., :3], a))
