import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from PIL import Image
import oracle as O
from ocr_system_b200.image_preprocessing import ImagePreprocessor
O.build()
n = 32
pages = np.stack([O.synth_page(3508, 2480, s) for s in range(n)])
imgs = [Image.fromarray(pages[i]) for i in range(n)]
ip = ImagePreprocessor(max_dimension=960)
ip.preprocess_pages_for_azure(imgs[:4])
def T(label, f):
    torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize(); print(f"{label:28s} {(time.perf_counter()-t)*1e3:8.1f} ms"); return r
o = T("auto_orient x32", lambda: [ip.auto_orient(im) for im in imgs])
a = T("np.asarray x32", lambda: [np.asarray(im) for im in o])
s = T("np.stack", lambda: np.stack(a))
x = T("to device (pageable)", lambda: torch.from_numpy(s).to("cuda"))
y = T("preprocess_device", lambda: ip.preprocess_device(x)[0])
f = T("compress_pages_for_azure", lambda: ip.compress_pages_for_azure(y))
T("whole call", lambda: ip.preprocess_pages_for_azure(imgs))
