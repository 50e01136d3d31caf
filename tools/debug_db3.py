import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2, torch, ctypes as C
from oracle import db_post as D
from ocr_system_b200 import _abi
full = D.synth_prob_map(960, 960, 0, n_boxes=300)
x0, y0, x1, y1 = 430, 540, 480, 585
pred = np.full((960, 960), 0.05, np.float32); pred[y0:y1, x0:x1] = full[y0:y1, x0:x1]
sub = np.ascontiguousarray(pred[y0-20:y1+20, x0-20:x1+20])
Lb = _abi.lib(); n = 1; h, w = sub.shape; maxc = 16
x = torch.from_numpy(sub[None]).cuda()
boxes = torch.zeros((n, maxc, 4, 2), dtype=torch.int32, device='cuda'); scores = torch.zeros((n, maxc), device='cuda'); counts = torch.zeros(n, dtype=torch.int32, device='cuda')
wsb = Lb.lumina_db_workspace_bytes(n, h, w, maxc); ws = torch.zeros(wsb + 8*1024*1024, dtype=torch.uint8, device='cuda')
hw = np.array([[h, w]], np.int32)
_abi.check(Lb.lumina_db_postprocess(C.c_void_p(x.data_ptr()), n, h, w, float(np.float32(0.3)), 0.99, 1.5, maxc, 3, hw.ctypes.data_as(C.c_void_p), C.c_void_p(boxes.data_ptr()), C.c_void_p(scores.data_ptr()), C.c_void_p(counts.data_ptr()), C.c_void_p(ws.data_ptr()), wsb, None))
torch.cuda.synchronize()
def a256(v): return (v + 255) & ~255
px = h*w; off = 0
off = a256(off + n*px); off = a256(off + n*px*4); off = a256(off + n*maxc*4); off = a256(off + n*maxc*16); off = a256(off + n*8)
pool_off = off
wsn = ws.cpu().numpy()
lanes = wsn[pool_off + 8*100000: pool_off + 8*100000 + 8*32].view(np.int32).reshape(32, 2)[:, 0]
print("per-lane cnt", lanes.tolist(), "sum", lanes.sum())
