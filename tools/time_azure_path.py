"""The whole per-page job of the reference app (OCRService._process_single_image_sync up to the Azure call):
host A4 raster -> preprocess_for_azure -> JPEG bytes.  GPU: ImagePreprocessor.preprocess_pages_for_azure on a
batch of 64 pages (host PIL images in, bytes out, wall clock); CPU: the reference's call sequence
(oracle/reference_port.preprocess_for_azure) on the box's cores, one page per task.  Also asserts that both
produce the same files."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from PIL import Image
import oracle as O
from oracle import reference_port as RP
from ocr_system_b200.image_preprocessing import ImagePreprocessor

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    O.build()
    ImagePreprocessor()   # first, as in the app (the singleton exists before pages are rasterised): raises Pillow's block size
    pages = np.stack([O.synth_page(3508, 2480, s) for s in range(n)])
    imgs = [Image.fromarray(pages[i]) for i in range(n)]
    res = {}
    for md in (2000, 960):
        ip = ImagePreprocessor(max_dimension=md)
        ip.preprocess_pages_for_azure(imgs[:8])            # warm-up (plans, workspaces, pinned buffers)
        ip.preprocess_pages_for_azure(imgs)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); got = ip.preprocess_pages_for_azure(imgs); gpu_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        for im in imgs[:8]:
            ip.preprocess_for_azure(im)          # the app's per-page call (ocr_service.py:412-417)
        one_ms = (time.perf_counter() - t0) / 8 * 1e3
        cpu_n = min(n, 32)
        cpu_s, want = RP.run_pool_azure(pages[:cpu_n], md)
        same = sum(a == b for a, b in zip(got[:cpu_n], want))
        res[md] = {"pages": n, "gpu_ms_per_batch": round(gpu_s * 1e3, 1), "gpu_pages_per_s": round(n / gpu_s, 1),
                   "cpu_cores": os.cpu_count(), "cpu_pages": cpu_n, "cpu_pages_per_s": round(cpu_n / cpu_s, 2),
                   "identical_files": f"{same}/{cpu_n}",
                   "per_page_call_ms_gpu": round(one_ms, 1), "mean_kb": round(float(np.mean([len(b) for b in got])) / 1024, 1)}
        print(md, json.dumps(res[md]), flush=True)
    # single-core latency of the reference sequence, AFTER all pools: OpenCV used in the parent before a fork
    # deadlocks the forked workers
    for md in (2000, 960):
        t0 = time.perf_counter(); RP.preprocess_for_azure(pages[0], md); res[md]["per_page_call_ms_cpu_1core"] = round((time.perf_counter() - t0) * 1e3, 1)
        print(md, "cpu single call ms", res[md]["per_page_call_ms_cpu_1core"], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/azure_path_timing.json", "w"))

if __name__ == "__main__":
    main()
