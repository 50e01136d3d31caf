"""One DBPostProcess call on 256 maps 960x960 (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ocr_system_b200 import ops
from oracle import db_post as D
maps = np.stack([D.synth_prob_map(960, 960, s, n_boxes=500) for s in range(16)])
pred = torch.from_numpy(maps).cuda().repeat(16, 1, 1).contiguous()
src = np.tile(np.array([[960, 960]], np.int32), (256, 1))
for _ in range(2):
    boxes, scores, counts = ops.db_postprocess(pred, src, 0.3, 0.6, 1.5, 1000, 3)
torch.cuda.synchronize()
print(counts[:4].tolist())
