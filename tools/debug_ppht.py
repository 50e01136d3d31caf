import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
import oracle as O
from ocr_system_b200 import ops, _abi
L = _abi.lib()
pg = O.synth_page(3508, 2480, 0)
tw, th = O.target_size(2480, 3508, 960)
img = O.resize_lanczos(pg, tw, th)
edges = O.canny(O.gray_cv(img))
h, w = edges.shape
ref = O.ppht(edges)
x = torch.from_numpy(edges[None]).cuda()
n = 1
lines = torch.zeros((n, 4096, 4), dtype=torch.int32, device='cuda'); nl = torch.zeros(n, dtype=torch.int32, device='cuda')
wsb = L.lumina_ppht_workspace_bytes(n, h, w, 1.0, np.pi/180)
ws = torch.zeros(wsb, dtype=torch.uint8, device='cuda')
_abi.check(L.lumina_ppht(C.c_void_p(x.data_ptr()), n, h, w, 1.0, float(np.pi/180), 100, 100, 10, C.c_void_p(lines.data_ptr()), C.c_void_p(nl.data_ptr()), 4096, C.c_void_p(ws.data_ptr()), wsb, None))
torch.cuda.synchronize()
def a256(v): return (v + 255) & ~255
numangle, numrho = 180, 2*(w+h)+1
px = h*w
accw = (numangle*numrho + 1)//2
off = 0
acc_off = off; off = a256(off + n*accw*4)
mask_off = off; off = a256(off + n*px)
nz_off = off; off = a256(off + n*px*4)
order_off = off; off = a256(off + n*px*4)
count_off = off; off = a256(off + n*4)
off = a256(off + numangle*2*4); off = a256(off + numangle*3*4); stats_off = off
wsn = ws.cpu().numpy()
N = wsn[count_off:count_off+4].view(np.int32)[0]
print("N", N, "ref N", int((edges>0).sum()))
order = wsn[order_off:order_off+4*N].view(np.uint32)
ys, xs = np.nonzero(edges)
nz = (ys.astype(np.uint32) << 16) | xs.astype(np.uint32)
# sequential order
state = (1<<64)-1; cnt = len(nz); arr = nz.copy(); ro = np.zeros(len(nz), np.uint32)
for i in range(len(nz)):
    state = ((state & 0xffffffff)*4164903690 + (state>>32)) & ((1<<64)-1)
    idx = (state & 0xffffffff) % cnt
    ro[i] = arr[idx]; arr[idx] = arr[cnt-1]; cnt -= 1
print("order equal", np.array_equal(order, ro), "first diff", (np.nonzero(order != ro)[0][:5] if not np.array_equal(order, ro) else None))
g = lines[0, :nl[0]].cpu().numpy()
print("nl", nl.item(), len(ref))
m = min(len(g), len(ref))
d = np.nonzero((g[:m] != ref[:m]).any(1))[0]
print("first differing line idx", d[:5])
if len(d):
    k = d[0]; print("gpu", g[max(0,k-1):k+3]); print("ref", ref[max(0,k-1):k+3])

print("stats N,votes,events,good,windows,batches,CS:", wsn[stats_off:stats_off+32].view(np.int32))
print("phase cycles fill,vote,reduce,sync1,rollback,walk,sync2,unvote,sync3,init:", wsn[stats_off+32:stats_off+32+80].view(np.int64))
import time
for nn in (1, 8, 64):
    xs_ = x.expand(nn, h, w).contiguous()
    for it in range(3):
        torch.cuda.synchronize(); t=time.time(); ops.hough_lines_p(xs_); torch.cuda.synchronize(); dt=time.time()-t
    print("n", nn, "ppht ms", dt*1e3)
