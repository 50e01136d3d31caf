#!/usr/bin/env python
"""bench.py -- A4 300-dpi pages/s of the page-image hot path on N B200s.

    python bench.py --gpus N --steps K --warmup W              (ours; N>1 under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   (CPU arm, rank 0 only)

A step = one pass of the chain (resize->960, deskew, gray, adaptive binarize, det
normalize: BASELINE.json configs[1]) over one batch of 64 synthetic A4 300-dpi RGB pages
per GPU.  `value` is timed with the rasters already resident in HBM; `e2e` goes through the
public host-buffer API (pinned host rasters -> HBM -> chain -> results back in host memory).
Pages shard across ranks with no data-path collective (weak scaling: 64 pages per GPU).
Timing: CUDA events on the launch stream, barrier + synchronize on both sides, MAX over ranks.
The input batch (1.67 GB) is larger than the 126 MB L2, so every step streams from HBM.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

PAGE_H, PAGE_W = 3508, 2480  # A4 @ 300 dpi
METRIC = "pages_per_sec"
UNIT = "pages/s"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                     0x4: "sw_power_cap"}
            getr = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(
                nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
            while not self._halt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = getr(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.1)   # NVML calls share a driver lock with kernel launches: sample sparsely
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def _physical_index(local: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


# ----------------------------------------------------------------------------- CPU arm
def _gen_page(seed):
    import oracle as O

    return O.synth_page(PAGE_H, PAGE_W, seed)


def host_pages(n: int):
    """n distinct synthetic pages generated on the host cores (oracle build of lumina_synth.h)."""
    import multiprocessing as mp

    import oracle as O

    O.build()
    with mp.get_context("fork").Pool(min(n, os.cpu_count() or 1)) as pool:
        return pool.map(_gen_page, range(n))


def cpu_reference_rate(sample_pages: int, max_dim: int, steps: int = 1, warmup: int = 0):
    """Reference CPU path (oracle/reference_port.py: the reference's Pillow/OpenCV calls) over
    a bounded sample, one page per task on every host core.  `sample_pages` page-tasks are drawn
    from min(sample_pages, 2*cores) distinct synthetic pages.  Returns (pages/s, s/step, cores)."""
    from oracle import reference_port as RP

    cores = os.cpu_count() or 1
    distinct = host_pages(min(sample_pages, 2 * cores))
    pages = [distinct[i % len(distinct)] for i in range(sample_pages)]
    for _ in range(warmup):
        RP.run_pool(distinct, max_dim, False, cores)
    ts = []
    for _ in range(steps):
        dt, _angles = RP.run_pool(pages, max_dim, False, cores)
        ts.append(dt)
    dt = sum(ts) / len(ts)
    return sample_pages / dt, dt, cores


def _workload(args):
    return {
        "workload": f"synthetic A4 300-dpi pages ({PAGE_W}x{PAGE_H} RGB) batch {args.batch} per GPU: "
                    f"resize-to-{args.max_dim} (PIL Lanczos) / deskew (Canny+HoughLinesP+bicubic warp) / "
                    "gray / adaptive binarize / det normalize  [BASELINE.json configs[1]]",
        "batch_per_gpu": args.batch, "max_dimension": args.max_dim,
        "cache": "inputs larger than L2 (1.67 GB batch vs 126 MB L2); no flush needed",
        "parallelism": "pages sharded by rank, no collective on the data path",
        "step_overlap": "value: consecutive steps are software-pipelined on CUDA streams (the wide kernels of step i+1 run "
                        "on the SMs that step i's last HoughLinesP clusters leave idle); value_unpipelined, stages_ms, "
                        "roofline and latency_bound come from the same K steps run one after the other",
    }


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample = args.cpu_sample or 8 * (os.cpu_count() or 1)
    rate, dt, cores = cpu_reference_rate(sample, args.max_dim, steps=args.steps, warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": _workload(args),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} synthetic A4 pages per step, one page per task, "
                                   f"multiprocessing.Pool({cores}), cv2.setNumThreads(1); "
                                   "oracle/reference_port.py = the reference's Pillow/OpenCV call sequence"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------- GPU arm
def main_ours(args):
    import torch
    import torch.distributed as dist

    from ocr_system_b200 import ops
    from ocr_system_b200.pipeline import PagePipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from ocr_system_b200.pipeline import bind_host_to_gpu_numa_node
    numa_node = bind_host_to_gpu_numa_node(local) if world > 1 else None   # staging memory next to this rank's GPU
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B, K, W = args.batch, args.steps, args.warmup
    pipe = PagePipeline(max_dimension=args.max_dim, device=dev)
    pages = ops.synth_pages(B, PAGE_H, PAGE_W, seed0=rank * B, device=dev)  # this rank's page range
    torch.cuda.synchronize()

    for _ in range(W):
        pipe.run_device(pages)
    barrier()

    # ---- timed region: device-resident input -------------------------------------
    sampler = ClockSampler(_physical_index(local))
    sampler.start()
    launches0 = ops.launch_count()
    timers = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        r = pipe.run_device(pages, profile=True)   # stage brackets = CUDA events on the launch stream
        timers.append(r.timer)
        angles = r.angles
        del r                                      # outputs are released every step (no growing pool)
    e1.record()
    barrier()
    unpipelined_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ops.launch_count() - launches0
    # ---- the headline `value`: the same K steps, software-pipelined on CUDA streams (PagePipeline.
    # run_device_stream): the wide kernels of step i+1 fill the SMs that step i's last HoughLinesP clusters leave
    # idle (49 pages fit at once, a 64-page batch alone runs 1.3 waves).  Same work, same results, every step's
    # outputs are produced and released; clock = barrier + event before the first step .. event + sync after the last.
    for r in pipe.run_device_stream([pages] * W):
        del r
    torch.cuda.synchronize()
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches_unpipelined, launches0 = launches, ops.launch_count()
    p0.record()
    for r in pipe.run_device_stream([pages] * K):
        angles = r.angles
        del r
    p1.record()
    torch.cuda.synchronize()
    barrier()
    elapsed_ms = max_over_ranks(p0.elapsed_time(p1))
    launches = ops.launch_count() - launches0      # kernels of the timed (pipelined) region
    clocks = sampler.stop()
    stage_ms = {}
    for t in timers:
        for k, v in t.collect().items():
            stage_ms[k] = stage_ms.get(k, 0.0) + v / K

    # ---- e2e: pinned host rasters -> HBM -> chain -> host results ------------------
    host_in = torch.empty(pages.shape, dtype=torch.uint8, pin_memory=True)
    host_in.copy_(pages)
    torch.cuda.synchronize()
    # PCIe probe (context for e2e): one pinned H2D copy of the batch
    _probe = torch.empty_like(pages)
    h2d_gbps = 0.0
    for _ in range(3):   # best of 3: the first copy pays the page-table / clock warm-up
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        _probe.copy_(host_in, non_blocking=True)
        h1.record()
        torch.cuda.synchronize()
        h2d_gbps = max(h2d_gbps, host_in.numel() / (h0.elapsed_time(h1) * 1e-3) / 1e9)
    del _probe
    # (a) cold: K batches through the public host-buffer API starting from an idle pipeline -- the first
    #     upload is not overlapped with anything, so this figure carries one pipeline fill per K steps.
    for _out, _res, h2d, d2h in pipe.run_host_stream([host_in] * 2):   # warm-up of the host path
        pass
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _out, _res, h2d, d2h in pipe.run_host_stream([host_in] * K):
        pass
    f1.record()
    barrier()
    e2e_cold_ms = max_over_ranks(f0.elapsed_time(f1))
    # (b) steady state (the headline e2e): ONE continuous stream of W + K + 1 batches.  The clock starts when
    #     the results of warm-up batch W-1 are in host memory and stops when those of batch W+K-1 are; the
    #     extra trailing batch keeps the upload pipe busy, so the region holds K uploads, K chains and K
    #     result downloads (shifted by one stage, as in any double-buffered stream).  Every batch's results
    #     are complete in host memory before the generator yields.
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    for i, (_out, _res, h2d, d2h) in enumerate(pipe.run_host_stream([host_in] * (W + K + 1))):
        if i == W - 1:
            g0.record()
        if i == W + K - 1:
            g1.record()
    torch.cuda.synchronize()
    barrier()
    e2e_ms = max_over_ranks(g0.elapsed_time(g1))

    if rank == 0:
        peak, peak_src = _peaks()
        ms_step = elapsed_ms / K
        value = world * B * K / (elapsed_ms / 1e3)
        e2e = world * B * K / (e2e_ms / 1e3)
        tw, th = ops.target_size(PAGE_W, PAGE_H, args.max_dim)
        # dominant HBM kernel = the fused Lanczos resize (reads the 26.1 MB raster once, writes the small page)
        alg_bytes = B * (PAGE_H * PAGE_W * 3 + th * tw * 3)
        rz_ms = stage_ms.get("resize_lanczos", float("nan"))
        achieved = alg_bytes / (rz_ms * 1e-3) / 1e9
        px = th * tw
        per_stage_bytes = {
            "resize_lanczos": alg_bytes, "canny": B * px * (3 + 1), "ppht": B * px,
            "angle+warp": B * px * 6, "gray_pil": B * px * 4, "adaptive_binarize": B * px * 2,
            "det_resize_normalize": B * (px * 3 + 3 * 960 * 672 * 4),
        }
        # DRAM bytes of one resize launch from the ncu --set full capture of this kernel at this batch size
        # (profiles/r1_ncu_resize_dp4a.txt: dram__bytes_read.sum 1.757 GB + dram__bytes_write.sum 71.1 MB)
        traffic, traffic_src = args.traffic_bytes, "--traffic-bytes"
        if traffic is None and B == 64 and args.max_dim == 960:
            traffic, traffic_src = 1.7570e9 + 71.1e6, "ncu capture profiles/r1_ncu_resize_dp4a.txt (batch 64, per launch)"
        elif traffic is None:
            traffic_src = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "value_unpipelined": world * B * K / (unpipelined_ms / 1e3), "ms_per_step_unpipelined": unpipelined_ms / K,
            "dtype": "u8", "data": "synthetic", "config": _workload(args),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / K, "pinned_h2d_GBps": round(h2d_gbps, 1),
                    "h2d_bound_ms_per_step": round(h2d / (h2d_gbps * 1e9) * 1e3, 2), "host_numa_node_rank0": numa_node,
                    "cold_start": {"value": world * B * K / (e2e_cold_ms / 1e3), "ms_per_step": e2e_cold_ms / K,
                                   "note": "same K batches from an idle pipeline (one un-overlapped upload per K steps)"},
                    "note": "steady state of PagePipeline.run_host_stream: pinned host rasters -> HBM -> chain -> results "
                            "in host memory for every batch; the upload of batch i+1 overlaps the kernels of batch i; "
                            "clock from 'results of warm-up batch W-1 in host memory' to 'results of batch W+K-1 in "
                            "host memory' inside one continuous stream of W+K+1 batches"},
            "gpu_launches": int(launches), "gpu_launches_unpipelined": int(launches_unpipelined),
            "clocks": clocks,
            "roofline": {
                "kernel": "resize_strip_dp4a_kernel<6> (fused PIL-Lanczos H+V, dominant HBM byte mover)",
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes,
            },
            "stages_ms": {k: round(v, 4) for k, v in stage_ms.items()},
            "stages_GBps": {k: round(per_stage_bytes[k] / (stage_ms[k] * 1e-3) / 1e9, 1)
                            for k in stage_ms if k in per_stage_bytes and stage_ms[k] > 0},
            "latency_bound": {"kernel": "ppht_cluster_pipe_kernel (exact cv2.HoughLinesP: serial dependency chain, "
                                        "3-CTA clusters, accumulator + edge bitmask in distributed shared memory)",
                              "ms_per_step": round(stage_ms.get("ppht", float("nan")), 3),
                              "share_of_step": round(stage_ms.get("ppht", 0.0) / ms_step, 3)},
            "deskew_angles_first4": [float(a) for a in angles[:4]],
        }
        if world == 1 and not args.no_cpu_baseline:
            sample = args.cpu_sample or 8 * (os.cpu_count() or 1)
            rate, dt, cores = cpu_reference_rate(sample, args.max_dim, steps=3, warmup=1)
            line["cpu_baseline"] = {
                "value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{sample} synthetic A4 page-tasks x 3 runs ({dt:.1f} s each, ~{dt * cores:.0f} core-s), one page per task on "
                          f"multiprocessing.Pool({cores}), cv2.setNumThreads(1); oracle/reference_port.py = "
                          "the reference's own Pillow/OpenCV call sequence",
            }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="pages per GPU per step")
    ap.add_argument("--max-dim", type=int, default=960)
    ap.add_argument("--cpu-sample", type=int, default=0, help="pages in the CPU baseline sample (default 2 x cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--traffic-bytes", type=float, default=None,
                    help="dram bytes/launch of the resize kernel from the committed ncu capture (profiles/)")
    args = ap.parse_args()
    if args.impl == "reference":
        return main_reference(args)
    args.warmup = max(args.warmup, 3)
    return main_ours(args)


if __name__ == "__main__":
    sys.exit(main())
