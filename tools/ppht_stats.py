import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from ocr_system_b200 import ops, _abi
L = _abi.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
pages = ops.synth_pages(n, 3508, 2480, 0)
small = ops.resize_if_needed(pages, 960)
edges = ops.canny(small)
h, w = edges.shape[1:]
lines = torch.zeros((n, 4096, 4), dtype=torch.int32, device='cuda'); nl = torch.zeros(n, dtype=torch.int32, device='cuda')
wsb = L.lumina_ppht_workspace_bytes(n, h, w, 1.0, np.pi/180)
ws = torch.zeros(wsb, dtype=torch.uint8, device='cuda')
_abi.check(L.lumina_ppht(C.c_void_p(edges.data_ptr()), n, h, w, 1.0, float(np.pi/180), 100, 100, 10, C.c_void_p(lines.data_ptr()), C.c_void_p(nl.data_ptr()), 4096, C.c_void_p(ws.data_ptr()), wsb, None))
torch.cuda.synchronize()
def a256(v): return (v + 255) & ~255
numangle, numrho = 180, 2*(w+h)+1
px = h*w; accw = (numangle*numrho + 1)//2
off = 0
off = a256(off + n*accw*4); off = a256(off + n*px); off = a256(off + n*px*4); off = a256(off + n*px*4); off = a256(off + n*4)
off = a256(off + numangle*2*4); off = a256(off + numangle*3*4); stats_off = off
wsn = ws.cpu().numpy()
st = wsn[stats_off:stats_off + n*32].view(np.int32).reshape(n, 8)
ll_off = stats_off + ((n*32 + 7) & ~7)
ph = wsn[ll_off: ll_off + n*80].view(np.int64).reshape(n, 10)
print("page  N     votes  events good batches | Mcycles: fill vote(sync) send wait events good+rollback rho groups")
for i in range(n):
    print(i, st[i, :4].tolist(), st[i, 5], "fast", st[i, 7], "|", (ph[i, :8] / 1e6).round(1).tolist(), "total", round(ph[i].sum()/1e6, 1))
