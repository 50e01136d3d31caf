"""Device timings of BASELINE.json configs[2] (DBPostProcess, 256 x 960x960 maps) and configs[3]
(CTC greedy decode, 1024 x 40 x 6625) with CUDA events; CPU oracle timed beside them on a sample."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ocr_system_b200 import ops
from oracle import db_post as D
import oracle as O

def timed(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))

res = {}
# ---- configs[2]: DB post on 256 maps (16 distinct synthetic maps tiled) ----
maps = np.stack([D.synth_prob_map(960, 960, s, n_boxes=500) for s in range(16)])
pred = torch.from_numpy(maps).cuda().repeat(16, 1, 1).contiguous()
src = np.tile(np.array([[960, 960]], np.int32), (256, 1))
ms = timed(lambda: ops.db_postprocess(pred, src, 0.3, 0.6, 1.5, 1000, 3))
boxes, scores, counts = ops.db_postprocess(pred, src, 0.3, 0.6, 1.5, 1000, 3)
t = time.perf_counter(); ref = D.DBPostProcess(0.3, 0.6, 1000, 1.5)({"maps": maps[:4, None]}, [(960, 960, 1.0, 1.0)] * 4); cpu = (time.perf_counter() - t) / 4
res["db_postprocess"] = {"maps": 256, "ms": ms, "maps_per_s": 256 / ms * 1e3, "boxes_per_map": float(counts.float().mean()),
                         "alg_GBps": 256 * 8.3e6 / ms / 1e6, "cpu_restated_ms_per_map": cpu * 1e3,
                         "ref_boxes_first4": [int(len(r["points"])) for r in ref], "gpu_boxes_first4": counts[:4].tolist()}
print(res["db_postprocess"], flush=True)
del pred
# ---- configs[3]: CTC greedy on 1024 x 40 x 6625 ----
g = torch.Generator(device="cuda").manual_seed(0)
p = torch.rand((1024, 40, 6625), device="cuda", generator=g) * 0.05
win = torch.randint(0, 6625, (1024, 40), device="cuda", generator=g)
p.scatter_(2, win[..., None], 0.9)
ms = timed(lambda: ops.ctc_greedy(p))
idx, pos, ln, conf = ops.ctc_greedy(p)
ph = p[:64].cpu().numpy()
t = time.perf_counter(); r = O.ctc_greedy(ph); cpu = time.perf_counter() - t
t = time.perf_counter(); am = ph.argmax(2); mx = ph.max(2); cpu_np = time.perf_counter() - t
ok = np.array_equal(idx[:64].cpu().numpy(), r[0]) and np.array_equal(ln[:64].cpu().numpy(), r[2])
res["ctc_greedy"] = {"shape": [1024, 40, 6625], "ms": ms, "GBps": p.numel() * 4 / ms / 1e6, "crops_per_s": 1024 / ms * 1e3,
                     "cpu_numpy_argmax_ms_per_1024": cpu_np * 16 * 1e3, "cpu_oracle_c_ms_per_1024": cpu * 16 * 1e3, "matches_oracle_64": bool(ok)}
print(res["ctc_greedy"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/paddle_ops_timing.json", "w"), indent=1)
