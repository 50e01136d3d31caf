"""JPEG row: 64 pages at 678x960 (and 1414x2000), quality 95, optimize=True (compress_for_azure's
first rung).  GPU: CUDA-event time of one encode call (device pages -> host files) and of the kernels only
(ncu-free: events around a reuse-free call); CPU: Pillow on the box's cores (multiprocessing, one page per task)."""
import io, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from PIL import Image
from ocr_system_b200 import ops

def pil_one(a):
    b = io.BytesIO(); Image.fromarray(a).save(b, format="JPEG", quality=95, optimize=True); return len(b.getvalue())

def main():
    import multiprocessing as mp
    res = {}
    for md in (960, 2000):
        n = 64
        pages = ops.synth_pages(n, 3508, 2480, 0)
        x = ops.resize_if_needed(pages, md)
        del pages
        x, _ = ops.deskew(x)
        x = ops.contrast_sharpness(x, 1.2, 1.1)
        h, w = x.shape[1], x.shape[2]
        enc = ops.JpegEncoder()
        files, sizes = enc.encode(x, 95, True)
        torch.cuda.synchronize()
        t = []
        for _ in range(5):
            t0 = time.perf_counter(); files, sizes = enc.encode(x, 95, True); t.append(time.perf_counter() - t0)
        gpu_ms = min(t) * 1e3
        host = x.cpu().numpy()
        cores = os.cpu_count()
        with mp.Pool(cores) as pool:
            pool.map(pil_one, [host[i] for i in range(min(n, cores))])
            t0 = time.perf_counter(); s = pool.map(pil_one, [host[i] for i in range(n)]); cpu_s = time.perf_counter() - t0
        assert [len(f) for f in files] == s
        t0 = time.perf_counter(); pil_one(host[0]); one = time.perf_counter() - t0
        res[md] = {"h": h, "w": w, "pages": n, "gpu_ms_per_batch": round(gpu_ms, 2), "gpu_pages_per_s": round(n / gpu_ms * 1e3),
                   "mean_file_kb": round(float(np.mean(s)) / 1024, 1), "cpu_cores": cores, "cpu_pool_pages_per_s": round(n / cpu_s, 1),
                   "cpu_ms_per_page_1core": round(one * 1e3, 1)}
        print(md, json.dumps(res[md]))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/jpeg_timing.json", "w"))

if __name__ == "__main__":
    main()
