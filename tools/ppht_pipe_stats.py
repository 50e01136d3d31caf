"""Per-phase cycle counts of the control warp of ppht_cluster_pipe_kernel (library built with LUMINA_PPHT_PROFILE=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from ocr_system_b200 import ops, _abi
L = _abi.lib()
L.lumina_ppht_stats_offset.restype = C.c_size_t
L.lumina_ppht_stats_offset.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
pages = ops.synth_pages(n, 3508, 2480, 0)
edges = ops.canny(ops.resize_if_needed(pages, 960))
h, w = edges.shape[1:]
lines = torch.zeros((n, 4096, 4), dtype=torch.int32, device='cuda'); nl = torch.zeros(n, dtype=torch.int32, device='cuda')
wsb = L.lumina_ppht_workspace_bytes(n, h, w, 1.0, np.pi / 180)
ws = torch.zeros(wsb, dtype=torch.uint8, device='cuda')
for _ in range(2):
    _abi.check(L.lumina_ppht(C.c_void_p(edges.data_ptr()), n, h, w, 1.0, float(np.pi / 180), 100, 100, 10, C.c_void_p(lines.data_ptr()),
                             C.c_void_p(nl.data_ptr()), 4096, C.c_void_p(ws.data_ptr()), wsb, None))
torch.cuda.synchronize()
off = int(L.lumina_ppht_stats_offset(n, h, w, 1.0, float(np.pi / 180)))
wsn = ws.cpu().numpy()
st = wsn[off:off + n * 32].view(np.int32).reshape(n, 8)
ll = off + ((n * 32 + 7) & ~7)
ph = wsn[ll: ll + n * 80].view(np.int64).reshape(n, 10)
print("page N flushes events lines steps | % of kernel: exch_wait events snapshot barrier_wait flush send | total Mcycles")
for i in range(n):
    t = ph[i, 9]
    print(i, st[i, 0], st[i, 1], st[i, 2], st[i, 3], st[i, 5], "|", " ".join(f"{100 * ph[i, j] / t:5.1f}" for j in (0, 1, 2, 3, 4, 5)), "|", round(t / 1e6, 2))
