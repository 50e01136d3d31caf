import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2, torch
from oracle import db_post as D
from ocr_system_b200.paddle_ops import DBPostProcess
pred = D.synth_prob_map(960, 960, 0, n_boxes=300)
sl = [(960, 960, 1.0, 1.0)]
kw = dict(thresh=0.3, box_thresh=0.6, unclip_ratio=1.5)
ref = D.DBPostProcess(**kw)({"maps": pred[None, None]}, sl, with_scores=True)[0]
got = DBPostProcess(**kw)({"maps": pred[None, None]}, sl, with_scores=True)[0]
rb, gb = ref["points"], got["points"]
print(len(rb), len(gb))
gs = set(tuple(b.reshape(-1)) for b in gb)
missing = [i for i, b in enumerate(rb) if tuple(b.reshape(-1)) not in gs]
print("missing idx", missing)
mask = (pred > np.float32(0.3)).astype(np.uint8)
cs, hier = cv2.findContours(mask * 255, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_SIMPLE)
print("contours", len(cs), "holes", int((hier[0][:, 3] != -1).sum()))
for i in missing[:6]:
    b = rb[i]; cx, cy = b[:, 0].mean(), b[:, 1].mean()
    # which contour contains this centre?
    for ci, c in enumerate(cs):
        if cv2.pointPolygonTest(c, (float(cx), float(cy)), False) >= 0:
            print(i, b.tolist(), "score", ref["scores"][i], "contour", ci, "is_hole", hier[0][ci][3] != -1, "npts", len(c))
rs = set(tuple(b.reshape(-1)) for b in rb)
extra = [i for i, b in enumerate(gb) if tuple(b.reshape(-1)) not in rs]
print("extra idx", extra[:10])

# ---- raw call keeping the workspace: per-slot reject codes ----
import ctypes as C
from ocr_system_b200 import _abi, ops
Lb = _abi.lib()
n, h, w, maxc = 1, 960, 960, 1000
x = torch.from_numpy(pred[None]).cuda()
boxes = torch.zeros((n, maxc, 4, 2), dtype=torch.int32, device='cuda'); scores = torch.zeros((n, maxc), device='cuda'); counts = torch.zeros(n, dtype=torch.int32, device='cuda')
wsb = Lb.lumina_db_workspace_bytes(n, h, w, maxc); ws = torch.zeros(wsb, dtype=torch.uint8, device='cuda')
hw = np.array([[960, 960]], np.int32)
_abi.check(Lb.lumina_db_postprocess(C.c_void_p(x.data_ptr()), n, h, w, float(np.float32(0.3)), 0.6, 1.5, maxc, 3, hw.ctypes.data_as(C.c_void_p), C.c_void_p(boxes.data_ptr()), C.c_void_p(scores.data_ptr()), C.c_void_p(counts.data_ptr()), C.c_void_p(ws.data_ptr()), wsb, None))
torch.cuda.synchronize()
def a256(v): return (v + 255) & ~255
px = h*w; off = 0
mask_off = off; off = a256(off + n*px)
labels_off = off; off = a256(off + n*px*4)
cand_off = off; off = a256(off + n*maxc*4)
bbox_off = off; off = a256(off + n*maxc*16)
ncand_off = off; off = a256(off + n*8)
pool = px + 4096
pool_off = off; off = a256(off + n*pool*8)
poolctr_off = off; off = a256(off + n*4)
accept_off = off; off = a256(off + n*maxc)
tmpbox_off = off; off = a256(off + n*maxc*32)
tmpscore_off = off; off = a256(off + n*maxc*4)
wsn = ws.cpu().numpy()
ncand = wsn[ncand_off:ncand_off+8].view(np.int32); print("ncand", ncand)
acc = wsn[accept_off:accept_off+ncand[1]]
cand = wsn[cand_off:cand_off+4*ncand[1]].view(np.int32)
bbox = wsn[bbox_off:bbox_off+16*ncand[1]].view(np.int32).reshape(-1,4)
tsc = wsn[tmpscore_off:tmpscore_off+4*ncand[1]].view(np.float32)
print("codes hist", np.unique(acc, return_counts=True))
csl, _ = cv2.findContours(mask*255, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
# per-contour oracle decision
k = 0
for slot in range(ncand[1]):
    c = csl[slot]
    points, sside = D.get_mini_boxes(c)
    ok_ref = False; sc = None
    if sside >= 3:
        sc = D.box_score_fast(pred, np.array(points).reshape(-1,2))
        if sc >= 0.6:
            ex = D.unclip(np.array(points), 1.5)
            if ex is not None and D.get_mini_boxes(ex.reshape(-1,1,2))[1] >= 5: ok_ref = True
    if ok_ref != (acc[slot] == 1) and k < 8:
        k += 1
        r = cand[slot]; r = -1-r if r < 0 else r
        tb = wsn[tmpbox_off + 32*slot: tmpbox_off + 32*slot + 32].view(np.int32)
        print("   gpu cnt,sum*1000,xmin,ymin,mw,mh,q0:", tb.tolist())
        print("slot", slot, "code", acc[slot], "root", (r % w, r // w), "contour first", c[0,0].tolist(), "bbox", bbox[slot].tolist(), "cv bbox", cv2.boundingRect(c), "gpu score", tsc[slot], "ref score", sc)
