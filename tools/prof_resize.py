import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ocr_system_b200 import ops
pages = ops.synth_pages(64, 3508, 2480, 0)
for _ in range(3):
    small = ops.resize_if_needed(pages, int(sys.argv[1]) if len(sys.argv) > 1 else 960)
torch.cuda.synchronize()
print(small.shape)
