"""One pass over the round-2 kernels for ncu (python tools/profile_round2.py): the chain on 64 A4 pages, the JPEG
decoder on the same pages as files, DBPostProcess on 256 maps, CTC, Otsu / Sauvola, the fast skew estimator."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocr_system_b200 import ops  # noqa: E402
from ocr_system_b200.pipeline import PagePipeline  # noqa: E402

torch.cuda.set_device(0)
pages = ops.synth_pages(64, 3508, 2480, seed0=0)
files = []
for i in range(0, 64, 16):
    files += ops.jpeg_encode(pages[i:i + 16], quality=75)
res = PagePipeline(max_dimension=960).run_device(pages)
dec = ops.JpegDecoder()
blob, offs = dec.pack(files)
out, status = dec.decode(blob, offs)
torch.cuda.synchronize()
assert int(status.abs().sum()) == 0
maps = ops.synth_prob_maps(256, 960, 960, seed0=0)
ops.db_postprocess(maps, [(960, 960)] * 256, 0.3, 0.6, 1.5, 1000)
post = ops.synth_ctc(1024, 40, 6625)
ops.ctc_greedy(post)
ops.otsu_binarize(res.gray)
ops.sauvola_binarize(res.gray, 25, 0.2, 128.0)
ops.estimate_skew_fast(ops.canny(res.pages, 50, 150))
torch.cuda.synchronize()
print("ok")
