#!/usr/bin/env python
"""bench.py -- A4 300-dpi pages/s of the page-image hot path on N B200s.

    python bench.py --gpus N --steps K --warmup W [--config 2|3|4|5|all]    (ours; N>1 under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W          (CPU arm, rank 0 only)

Workloads (BASELINE.json `configs`):
  2  (headline, default top-level line) 64 synthetic A4 300-dpi RGB pages per GPU per step through
     resize->960 / deskew / gray / adaptive binarize / det normalize.  `value`: rasters resident in HBM,
     consecutive steps stream-pipelined; `e2e`: pinned host rasters -> HBM -> chain -> results in host memory.
  3  DBPostProcess on 256 synthetic 960x960 probability maps (~500 boxes per map).
  4  CTC greedy decode of 1024 x 40 x 6625 posteriors.
  5  10 000-page stream sharded by page over the ranks: per 64-page batch the chain, then DBPostProcess on a
     synthetic detector map per page and CTC decode of synthetic posteriors (crops per page stated in the line).
With `--config all` (the default) the top-level JSON line is config 2 and the same line carries
`"other_configs": {"3": {...}, "4": {...}, "5": {...}}`, each measured in the same process.

A "step" = one pass of the hot path over one batch.  Every timed region rotates over >= 4 distinct resident
batches and repeats the K-step round until >= ~2 s have been timed (`rounds`, `timed_steps` in the line).
Timing: CUDA events on the launch stream, barrier + synchronize on both sides, MAX over ranks.  All batches
are larger than the 126 MB L2, so every step streams from HBM.
"""
from __future__ import annotations

import argparse
import json
import math
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
MIN_TIMED_S = 2.0
DB_KW = dict(thresh=0.3, box_thresh=0.6, unclip_ratio=1.5, max_candidates=1000)
CTC_T, CTC_C = 40, 6625
JPEG_Q = 75   # Pillow's default quality: the file form of a page for e2e_compressed


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _ncu_traffic(key: str):
    """dram bytes per launch of the named kernel from the committed ncu --set full capture (profiles/)."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            e = json.load(f)[key]
        return float(e["dram_bytes_per_launch"]), e["source"]
    except Exception:
        return None, None


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


# ============================================================================= CPU arm (oracle side)
def _gen_page(seed):
    import oracle as O

    return O.synth_page(PAGE_H, PAGE_W, seed)


def host_pages(n: int, seed0: int = 0):
    """n distinct synthetic pages generated on the host cores (oracle build of lumina_synth.h)."""
    import multiprocessing as mp

    import oracle as O

    O.build()
    with mp.get_context("fork").Pool(min(n, os.cpu_count() or 1)) as pool:
        return pool.map(_gen_page, range(seed0, seed0 + n))


def _enc_page(page):
    import io

    from PIL import Image

    b = io.BytesIO()
    Image.fromarray(page).save(b, "JPEG", quality=JPEG_Q)
    return b.getvalue()


def cpu_reference_rate_from_files(sample_pages: int, max_dim: int, steps: int = 1):
    """Same as cpu_reference_rate, but every page-task starts from the page's JPEG file (quality JPEG_Q) and calls
    the reference's load_image_bytes first -- the CPU side of e2e_compressed."""
    import multiprocessing as mp

    from oracle import reference_port as RP

    cores = os.cpu_count() or 1
    distinct = host_pages(min(sample_pages, 2 * cores))
    with mp.get_context("fork").Pool(min(len(distinct), cores)) as pool:
        files = pool.map(_enc_page, distinct)
    tasks = [files[i % len(files)] for i in range(sample_pages)]
    ts = [RP.run_pool(tasks, max_dim, False, cores)[0] for _ in range(steps)]
    dt = sum(ts) / len(ts)
    return sample_pages / dt, dt, cores


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


_CPU_SHARED = {}


def _cpu_db_task(i):
    from oracle import db_post as D

    m = _CPU_SHARED["maps"][i]
    r = D.DBPostProcess(**DB_KW)({"maps": m[None, None]}, [(m.shape[0], m.shape[1], 1.0, 1.0)])
    return len(r[0]["points"])


def _cpu_ctc_task(i):
    """upstream CTCLabelDecode.__call__ on one slab of crops: NumPy argmax / max, Python collapse + join."""
    import numpy as np

    p = _CPU_SHARED["post"][i]
    chars = _CPU_SHARED["chars"]
    idx, prob = p.argmax(axis=2), p.max(axis=2)
    out = []
    for b in range(idx.shape[0]):
        sel = np.ones(idx.shape[1], bool)
        sel[1:] = idx[b, 1:] != idx[b, :-1]
        sel &= idx[b] != 0
        out.append(("".join(chars[k] for k in idx[b][sel]), float(np.mean(prob[b][sel])) if sel.any() else 0.0))
    return len(out)


def _cpu_pool_rate(task, n_tasks: int, units_per_task: int, repeats: int):
    import multiprocessing as mp

    cores = os.cpu_count() or 1
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(task, range(min(n_tasks, cores)))              # warm the workers
        ts = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            pool.map(task, range(n_tasks), chunksize=1)
            ts.append(time.perf_counter() - t0)
    dt = sum(ts) / len(ts)
    return n_tasks * units_per_task / dt, dt, cores


def cpu_db_rate(h: int, w: int, repeats: int = 3):
    """oracle/db_post.py (upstream DBPostProcess restated with cv2 + restated Clipper) on 8 maps per core."""
    import oracle as O

    cores = os.cpu_count() or 1
    n = 8 * cores
    _CPU_SHARED["maps"] = [O.synth_prob_map_grid(h, w, s) for s in range(n)]
    rate, dt, cores = _cpu_pool_rate(_cpu_db_task, n, 1, repeats)
    return rate, dt, cores, f"{n} synthetic {w}x{h} maps x {repeats} runs ({dt:.2f} s each), one map per task"


def _ctc_chars():
    """6623 dictionary entries (Devanagari block with combining marks first) + space: 6625 classes with the blank."""
    chars = [chr(c) for c in range(0x0900, 0x0980)]
    c = 0x4E00
    while len(chars) < CTC_C - 2:
        chars.append(chr(c))
        c += 1
    return ["blank"] + chars + [" "]


def cpu_ctc_rate(repeats: int = 10, crops_per_task: int = 32):
    import oracle as O

    cores = os.cpu_count() or 1
    n = 4 * cores
    _CPU_SHARED["post"] = [O.synth_ctc(crops_per_task, CTC_T, CTC_C, i * crops_per_task, 1) for i in range(n)]
    _CPU_SHARED["chars"] = _ctc_chars()
    rate, dt, cores = _cpu_pool_rate(_cpu_ctc_task, n, crops_per_task, repeats)
    return rate, dt, cores, (f"{n * crops_per_task} crops x {CTC_T} x {CTC_C} x {repeats} runs ({dt:.2f} s each), "
                             f"{crops_per_task} crops per task, NumPy argmax/max + Python collapse (upstream CTCLabelDecode)")


def _chain_kind():
    from oracle import reference_port as RP

    return RP.kind()


def _chain_kind_note():
    if _chain_kind() == "reference":
        return ("oracle/_ref/image_preprocessing.py = the reference's own module (unmodified copy placed by "
                "oracle/make_ref.py), its ImagePreprocessor methods called in the bench chain's order")
    return "oracle/reference_port.py = the reference's Pillow/OpenCV call sequence"


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = "2" if args.config == "all" else args.config
    cores = os.cpu_count() or 1
    base = {"impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "data": "synthetic"}

    def chain_line():
        sample = args.cpu_sample or 8 * cores
        rate, dt, _ = cpu_reference_rate(sample, args.max_dim, steps=args.steps, warmup=min(args.warmup, 1))
        frate, _fdt, _ = cpu_reference_rate_from_files(sample, args.max_dim, steps=max(1, min(args.steps, 3)))
        base["e2e_compressed"] = {"value": frate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                                  "note": f"same page-tasks starting from the page's JPEG file (quality {JPEG_Q}): the "
                                          "reference's load_image_bytes (Pillow decode) + the chain"}
        return dict(base, metric=METRIC, value=rate, unit=UNIT, ms_per_step=dt * 1e3, dtype="u8", config=_workload2(args),
                    cpu_baseline={"value": rate, "unit": UNIT, "cores": cores, "kind": _chain_kind(),
                                  "sample": f"{sample} synthetic A4 pages per step, one page per task, "
                                            f"multiprocessing.Pool({cores}), cv2.setNumThreads(1); " + _chain_kind_note()},
                    e2e={"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})

    def db_line():
        rate, dt, _, sample = cpu_db_rate(960, 960, repeats=max(1, min(args.steps, 5)))
        return dict(base, metric=METRIC, value=rate, unit="maps/s", ms_per_step=dt * 1e3, dtype="f32", config=_workload3(256),
                    cpu_baseline={"value": rate, "unit": "maps/s", "cores": cores, "kind": "port", "sample": sample},
                    e2e={"value": rate, "unit": "maps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})

    def ctc_line():
        rate, dt, _, sample = cpu_ctc_rate(repeats=max(1, min(args.steps, 5)))
        return dict(base, metric=METRIC, value=rate, unit="crops/s", ms_per_step=dt * 1e3, dtype="f32", config=_workload4(1024),
                    cpu_baseline={"value": rate, "unit": "crops/s", "cores": cores, "kind": "port", "sample": sample},
                    e2e={"value": rate, "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})

    def stream_line():
        # pages/s of chain + DB + CTC per page = harmonic combination of the three CPU rates (independent stages
        # on the same cores), measured on bounded samples of each
        sample = args.cpu_sample or 4 * cores
        r2, _, _ = cpu_reference_rate(sample, args.max_dim, steps=1, warmup=0)
        r3, _, _, s3 = cpu_db_rate(960, 672, repeats=1)
        r4, _, _, s4 = cpu_ctc_rate(repeats=1)
        cpp = args.crops_per_page
        rate = 1.0 / (1.0 / r2 + 1.0 / r3 + cpp / r4)
        return dict(base, metric=METRIC, value=rate, unit=UNIT, ms_per_step=64e3 / rate, dtype="u8",
                    config=_workload5(args, args.stream_pages),
                    cpu_baseline={"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                  "sample": f"(chain leg: kind {_chain_kind()}) extrapolated from bounded samples: chain {r2:.1f} pages/s on {sample} pages; "
                                            f"DB {r3:.1f} maps/s ({s3}); CTC {r4:.0f} crops/s ({s4}); {cpp} crops per page"},
                    e2e={"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})

    line = {"2": chain_line, "3": db_line, "4": ctc_line, "5": stream_line}[cfg]()
    print(json.dumps(line), flush=True)
    return 0


# ============================================================================= workload descriptions
def _workload2(args):
    return {
        "workload": f"synthetic A4 300-dpi pages ({PAGE_W}x{PAGE_H} RGB) batch {args.batch} per GPU: "
                    f"resize-to-{args.max_dim} (PIL Lanczos) / deskew (Canny+HoughLinesP+bicubic warp) / "
                    "gray / adaptive binarize / det normalize  [BASELINE.json configs[1]]",
        "batch_per_gpu": args.batch, "max_dimension": args.max_dim,
        "cache": "inputs larger than L2 (1.67 GB batch vs 126 MB L2); no flush needed",
        "parallelism": "pages sharded by rank, no collective on the data path",
        "step_overlap": "value: consecutive steps are software-pipelined on CUDA streams (the wide kernels of step i+1 run "
                        "on the SMs that step i's last HoughLinesP clusters leave idle); value_unpipelined, stages_ms, "
                        "roofline and latency_bound come from the same steps run one after the other",
    }


def _workload3(n_maps):
    return {"workload": f"DBPostProcess on synthetic 960x960 probability maps batch {n_maps} per GPU (~500 boxes/map, "
                        "thresh 0.3, box_thresh 0.6, unclip 1.5, max_candidates 1000, score_mode fast) + box score/unclip  "
                        "[BASELINE.json configs[2]]",
            "batch_per_gpu": n_maps, "cache": "inputs larger than L2 (944 MB batch vs 126 MB L2); no flush needed",
            "parallelism": "maps sharded by rank, no collective on the data path"}


def _workload4(n_crops):
    return {"workload": f"CTC greedy decode on synthetic posteriors batch {n_crops} crops x T={CTC_T} x C={CTC_C} per GPU "
                        "(6.6k-class dictionary incl. the Devanagari block; planted repeats, blanks, exact ties)  "
                        "[BASELINE.json configs[3]]",
            "batch_per_gpu": n_crops, "cache": "inputs larger than L2 (1.085 GB batch vs 126 MB L2); no flush needed",
            "parallelism": "crops sharded by rank, no collective on the data path"}


def _workload5(args, n_pages):
    return {"workload": f"synthetic {n_pages}-page raster stream sharded by page across the GPUs, batches of {args.batch}: "
                        f"preprocess chain (resize-to-{args.max_dim}, deskew, gray, adaptive binarize, det normalize) + "
                        f"DBPostProcess on a synthetic {'672x960'} detector map per page + CTC greedy decode of "
                        f"{args.crops_per_page} synthetic crops per page (T={CTC_T}, C={CTC_C})  [BASELINE.json configs[4]]",
            "pages_total": n_pages, "batch_per_gpu": args.batch, "crops_per_page": args.crops_per_page,
            "cache": "two distinct resident batches (rasters 1.67 GB + detector maps + posteriors 34 GB each, far larger than L2) "
                     "rotated through the stream",
            "parallelism": "contiguous page ranges per rank (shard_range), no collective on the data path",
            "timing": "value = pages / CUDA-event time of the rank's whole stream from an idle pipeline (max over ranks): the chains "
                      "of consecutive batches are software-pipelined (PagePipeline.run_device_stream) and DB + CTC of batch i run on "
                      "a side stream beside the chain of batch i+1; the synthetic generators stand in for the rasteriser / detector / "
                      "recogniser networks and run before the timed region"}


# ============================================================================= GPU arm
class Ctx:
    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.args = args

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v: float) -> float:
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v: float) -> float:
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def timed_rounds(self, one_round, K: int, est_ms_per_step: float):
        """Run `one_round()` (exactly K steps) R times, R chosen so that >= MIN_TIMED_S are timed; each round is
        bracketed by barrier + synchronize and CUDA events; returns (total ms as MAX over ranks per round summed, R)."""
        torch = self.torch
        R = max(1, int(math.ceil(MIN_TIMED_S * 1e3 / max(est_ms_per_step * K, 1e-3))))
        R = min(R, 4096)
        if self.world > 1:   # every rank must run the same number of rounds
            R = int(self.max_over_ranks(float(R)))
        total = 0.0
        for _ in range(R):
            self.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            one_round()
            b.record()
            self.barrier()
            total += self.max_over_ranks(a.elapsed_time(b))
        return total, R


def bench_chain(cx: Ctx):
    """config 2 -> the top-level line."""
    import numpy as np

    torch = cx.torch
    args = cx.args
    from ocr_system_b200 import ops
    from ocr_system_b200.pipeline import PagePipeline

    B, K, W, P = args.batch, args.steps, args.warmup, args.pool
    dev, rank, world = cx.dev, cx.rank, cx.world
    pipe = PagePipeline(max_dimension=args.max_dim, device=dev)
    # P distinct resident batches per rank (page seeds never repeat across ranks or batches)
    pool = [ops.synth_pages(B, PAGE_H, PAGE_W, seed0=(rank * P + p) * B, device=dev) for p in range(P)]
    torch.cuda.synchronize()

    for i in range(W):
        pipe.run_device(pool[i % P])
    cx.barrier()

    sampler = ClockSampler(_physical_index(cx.local))
    sampler.start()
    # ---- (a) steps one after the other: stage brackets, roofline, launch count per step ---------------------
    timers = []
    state = {"angles": None, "i": 0}

    def round_unpipelined():
        for _ in range(K):
            r = pipe.run_device(pool[state["i"] % P], profile=True)
            state["i"] += 1
            timers.append(r.timer)
            state["angles"] = r.angles
            del r

    # estimate one step for the round count
    t0 = time.perf_counter()
    pipe.run_device(pool[0])
    torch.cuda.synchronize()
    est = (time.perf_counter() - t0) * 1e3
    launches0 = ops.launch_count()
    unp_ms, R_unp = cx.timed_rounds(round_unpipelined, K, est)
    launches_unp = (ops.launch_count() - launches0) // (R_unp * K)
    stage_ms = {}
    for t in timers:
        for k, v in t.collect().items():
            stage_ms[k] = stage_ms.get(k, 0.0) + v / len(timers)
    timers.clear()

    # ---- (b) the headline `value`: the same steps, software-pipelined on CUDA streams ----------------------
    for r in pipe.run_device_stream([pool[i % P] for i in range(W)]):
        del r
    torch.cuda.synchronize()

    def round_pipelined():
        s = state["i"]
        for r in pipe.run_device_stream([pool[(s + i) % P] for i in range(K)]):
            state["angles"] = r.angles
            del r
        state["i"] += K

    launches0 = ops.launch_count()
    # the pipelined step is ~0.7 of the unpipelined one: estimate with 0.6 so that the timed region stays >= MIN_TIMED_S
    pip_ms, R_pip = cx.timed_rounds(round_pipelined, K, 0.6 * unp_ms / (R_unp * K))
    launches = ops.launch_count() - launches0
    clocks = sampler.stop()

    # ---- (b') the same steps with the flag-gated fast skew estimator instead of the exact HoughLinesP replica
    # (NOT the reference's angle: reported beside `value`, never as it) ----------------------------------------
    fpipe = PagePipeline(max_dimension=args.max_dim, device=dev, deskew_mode="fast")
    for r in fpipe.run_device_stream([pool[i % P] for i in range(W)]):
        del r
    torch.cuda.synchronize()

    def round_fast():
        s = state["i"]
        for r in fpipe.run_device_stream([pool[(s + i) % P] for i in range(K)]):
            del r
        state["i"] += K

    fast_ms, R_fast = cx.timed_rounds(round_fast, K, pip_ms / (R_pip * K) / 2)

    # ---- (c) e2e: pinned host rasters -> HBM -> chain -> host results --------------------------------------
    host_pool = []
    for p in range(P):
        hb = torch.empty(pool[p].shape, dtype=torch.uint8, pin_memory=True)
        hb.copy_(pool[p])
        host_pool.append(hb)
    torch.cuda.synchronize()
    _probe = torch.empty_like(pool[0])
    h2d_gbps = 0.0
    for _ in range(3):   # PCIe probe (context for e2e): best of 3 pinned H2D copies of one batch
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        _probe.copy_(host_pool[0], non_blocking=True)
        h1.record()
        torch.cuda.synchronize()
        h2d_gbps = max(h2d_gbps, host_pool[0].numel() / (h0.elapsed_time(h1) * 1e-3) / 1e9)
    del _probe
    for _out, _res, h2d, d2h in pipe.run_host_stream([host_pool[i % P] for i in range(2)]):   # warm-up of the host path
        pass
    # cold: K batches from an idle pipeline (one un-overlapped upload per K steps)
    cx.barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _out, _res, h2d, d2h in pipe.run_host_stream([host_pool[i % P] for i in range(K)]):
        pass
    f1.record()
    cx.barrier()
    e2e_cold_ms = cx.max_over_ranks(f0.elapsed_time(f1))
    # steady state (the headline e2e): ONE continuous stream of W + Ke + 1 batches; the clock starts when the results
    # of warm-up batch W-1 are in host memory and stops when those of batch W+Ke-1 are.  Ke = K * rounds (>= ~2 s).
    Re = max(1, min(64, int(math.ceil(MIN_TIMED_S * 1e3 / max(e2e_cold_ms, 1e-3)))))
    Ke = K * Re
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cx.barrier()
    for i, (_out, _res, h2d, d2h) in enumerate(pipe.run_host_stream(host_pool[i % P] for i in range(W + Ke + 1))):
        if i == W - 1:
            g0.record()
        if i == W + Ke - 1:
            g1.record()
    torch.cuda.synchronize()
    cx.barrier()
    e2e_ms = cx.max_over_ranks(g0.elapsed_time(g1))
    del host_pool

    # ---- (d) e2e_compressed: the pages start as baseline-JPEG FILES in pinned host memory (what load_image_bytes is
    # handed); files -> HBM -> device decode -> stream-pipelined chain -> results in host memory ---------------
    enc_pool, file_bytes = [], 0
    packer = ops.JpegDecoder()
    for p in range(P):
        files = []
        for j in range(0, B, 16):   # the device encoder writes Pillow's byte stream (tests/test_jpeg.py)
            files += ops.jpeg_encode(pool[p][j:j + 16], quality=JPEG_Q)
        blob, offs = packer.pack(files)
        enc_pool.append((blob.clone().pin_memory(), offs))
        file_bytes += int(offs[-1])
    del packer
    torch.cuda.synchronize()
    for _ in pipe.run_encoded_stream([enc_pool[i % P] for i in range(3)]):
        pass
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cx.barrier()
    launches0 = ops.launch_count()
    for i, (_out, _res, _h2d, d2h_c) in enumerate(pipe.run_encoded_stream(enc_pool[i % P] for i in range(W + Ke + 2))):
        if i == W - 1:
            c0.record()
        if i == W + Ke - 1:
            c1.record()
    torch.cuda.synchronize()
    cx.barrier()
    e2ec_ms = cx.max_over_ranks(c0.elapsed_time(c1))
    launches_c = (ops.launch_count() - launches0) // (W + Ke + 2)
    del enc_pool

    if rank != 0:
        return None
    peak, peak_src = _peaks()
    steps_timed = R_pip * K
    ms_step = pip_ms / steps_timed
    value = world * B * steps_timed / (pip_ms / 1e3)
    e2e = world * B * Ke / (e2e_ms / 1e3)
    tw, th = ops.target_size(PAGE_W, PAGE_H, args.max_dim)
    alg_bytes = B * (PAGE_H * PAGE_W * 3 + th * tw * 3)
    rz_ms = stage_ms.get("resize_lanczos", float("nan"))
    achieved = alg_bytes / (rz_ms * 1e-3) / 1e9
    px = th * tw
    per_stage_bytes = {
        "resize_lanczos": alg_bytes, "canny": B * px * (3 + 1), "ppht": B * px,
        "angle+warp": B * px * 6, "gray_pil": B * px * 4, "adaptive_binarize": B * px * 2,
        "det_resize_normalize": B * (px * 3 + 3 * 960 * 672 * 4),
    }
    traffic, traffic_src = args.traffic_bytes, "--traffic-bytes"
    if traffic is None:
        traffic, traffic_src = _ncu_traffic(f"resize@{B}x{PAGE_H}x{PAGE_W}->{args.max_dim}")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "rounds": R_pip, "timed_steps": steps_timed, "timed_region_s": pip_ms / 1e3, "distinct_batches": P,
        "value_unpipelined": world * B * R_unp * K / (unp_ms / 1e3), "ms_per_step_unpipelined": unp_ms / (R_unp * K),
        "value_fast_deskew": {"value": world * B * R_fast * K / (fast_ms / 1e3), "unit": UNIT, "ms_per_step": fast_ms / (R_fast * K),
                              "note": "deskew_mode='fast' (projection-profile angle, flag-gated, NOT the reference's "
                                      "HoughLinesP angle: rasters differ; profiles/r2_fast_skew_certification.json)"},
        "dtype": "u8", "data": "synthetic", "config": _workload2(args),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_ms / Ke, "timed_steps": Ke, "pinned_h2d_GBps": round(h2d_gbps, 1),
                "h2d_bound_ms_per_step": round(h2d / (h2d_gbps * 1e9) * 1e3, 2),
                "cold_start": {"value": world * B * K / (e2e_cold_ms / 1e3), "ms_per_step": e2e_cold_ms / K,
                               "note": "K batches from an idle pipeline (one un-overlapped upload per K steps)"},
                "note": "steady state of PagePipeline.run_host_stream over rotating distinct pinned batches: host rasters -> "
                        "HBM -> chain -> results in host memory for every batch; the upload of batch i+1 overlaps the "
                        "kernels of batch i"},
        "e2e_compressed": {"value": world * B * Ke / (e2ec_ms / 1e3), "unit": UNIT,
                           "h2d_bytes_per_step": int(file_bytes // P), "d2h_bytes_per_step": int(d2h_c),
                           "ms_per_step": e2ec_ms / Ke, "timed_steps": Ke, "gpu_launches_per_step": int(launches_c),
                           "note": f"steady state of PagePipeline.run_encoded_stream: the same pages as baseline-JPEG files "
                                   f"(quality {JPEG_Q}, 4:2:0, what Pillow's save() writes) in pinned host memory -> HBM -> "
                                   "device decode (raster == Pillow's, byte for byte) -> stream-pipelined chain -> results in "
                                   "host memory; the reference side of this is load_image_bytes + the chain"},
        "gpu_launches": int(launches), "gpu_launches_per_step": int(launches_unp),
        "clocks": clocks,
        "roofline": {
            "kernel": "resize_strip_bulk_kernel<2,2,true> (fused PIL-Lanczos H+V, TMA-staged, int8 mma.sync in both passes; the dominant HBM byte mover)",
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg_bytes,
        },
        "stages_ms": {k: round(v, 4) for k, v in stage_ms.items()},
        "stages_GBps": {k: round(per_stage_bytes[k] / (stage_ms[k] * 1e-3) / 1e9, 1)
                        for k in stage_ms if k in per_stage_bytes and stage_ms[k] > 0},
        "latency_bound": {"kernel": "HoughLinesP cluster kernel (exact cv2.HoughLinesP: serial dependency chain)",
                          "ms_per_step": round(stage_ms.get("ppht", float("nan")), 3),
                          "share_of_step": round(stage_ms.get("ppht", 0.0) / (unp_ms / (R_unp * K)), 3)},
        "deskew_angles_first4": [float(a) for a in np.asarray(state["angles"])[:4]],
    }
    if world == 1 and not args.no_cpu_baseline:
        sample = args.cpu_sample or 8 * (os.cpu_count() or 1)
        rate, dt, cores = cpu_reference_rate(sample, args.max_dim, steps=3, warmup=1)
        line["cpu_baseline"] = {
            "value": rate, "unit": UNIT, "cores": cores, "kind": _chain_kind(),
            "sample": f"{sample} synthetic A4 page-tasks x 3 runs ({dt:.1f} s each, ~{dt * cores:.0f} core-s), one page per task on "
                      f"multiprocessing.Pool({cores}), cv2.setNumThreads(1); " + _chain_kind_note(),
        }
    return line


def bench_db(cx: Ctx, top: bool):
    """config 3: DBPostProcess, 256 maps of 960x960 per GPU per step."""
    import numpy as np

    torch = cx.torch
    args = cx.args
    from ocr_system_b200 import ops
    from ocr_system_b200.paddle_ops import DBPostProcess

    N, H, Wd, P = args.db_maps, 960, 960, max(2, args.pool // 2)
    K, W = args.steps, args.warmup
    dev, rank, world = cx.dev, cx.rank, cx.world
    pool = [ops.synth_prob_maps(N, H, Wd, seed0=(rank * P + p) * N, device=dev) for p in range(P)]
    src = np.tile(np.array([[H, Wd]], np.int32), (N, 1))
    call = lambda pred: ops.db_postprocess(pred, src, DB_KW["thresh"], DB_KW["box_thresh"], DB_KW["unclip_ratio"],  # noqa: E731
                                           DB_KW["max_candidates"], 3)
    for i in range(W):
        call(pool[i % P])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    boxes, scores, counts = call(pool[0])
    torch.cuda.synchronize()
    est = (time.perf_counter() - t0) * 1e3
    mean_boxes = float(counts.float().mean())
    st = {"i": 0}

    def one_round():
        for _ in range(K):
            out = call(pool[st["i"] % P])
            st["i"] += 1
            del out

    sampler = ClockSampler(_physical_index(cx.local))
    sampler.start()
    l0 = ops.launch_count()
    ms, R = cx.timed_rounds(one_round, K, est)
    launches = ops.launch_count() - l0
    clocks = sampler.stop()
    # e2e through the upstream-signature operator: pinned host maps -> boxes / scores on the host
    host = torch.empty((N, 1, H, Wd), dtype=torch.float32, pin_memory=True)
    host[:, 0].copy_(pool[0])
    torch.cuda.synchronize()
    post = DBPostProcess(**DB_KW)
    shape_list = np.tile(np.array([[H, Wd, 1.0, 1.0]], np.float64), (N, 1))
    for _ in range(2):
        res = post({"maps": host}, shape_list, with_scores=True)
    Ke = max(3, min(K, 20))
    cx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(Ke):
        res = post({"maps": host}, shape_list, with_scores=True)
    e1.record()
    cx.barrier()
    e2e_ms = cx.max_over_ranks(e0.elapsed_time(e1))
    d2h = sum(r["points"].nbytes + r["scores"].nbytes // 2 for r in res) + 4 * N
    del host
    if rank != 0:
        return None
    peak, peak_src = _peaks()
    steps_timed = R * K
    per_map = 2 * H * Wd * 4 + 4 * 30 * 14 * mean_boxes   # prob read + labels written + score re-read ~ 4 x sum(bbox area)
    achieved = N * per_map / (ms / steps_timed * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": world * N * steps_timed / (ms / 1e3), "unit": "maps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / steps_timed, "rounds": R, "timed_steps": steps_timed, "timed_region_s": ms / 1e3,
        "distinct_batches": P, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": _workload3(N), "boxes_per_map": mean_boxes,
        "e2e": {"value": world * N * Ke / (e2e_ms / 1e3), "unit": "maps/s", "h2d_bytes_per_step": N * H * Wd * 4,
                "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / Ke, "timed_steps": Ke,
                "note": "paddle_ops.DBPostProcess(outs_dict{'maps': pinned host [N,1,H,W]}, shape_list): H2D of the maps, the "
                        "11 kernels, D2H of boxes / scores / counts, per-map Python dicts"},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"kernel": "lumina_db_postprocess (mask + labelling + candidate geometry + score + unclip)",
                     "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": N * per_map,
                     "note": "irregular per-box work; the labelling passes are atomic / latency bound, not bandwidth bound"},
    }
    if world == 1 and not args.no_cpu_baseline:
        rate, dt, cores, sample = cpu_db_rate(H, Wd)
        line["cpu_baseline"] = {"value": rate, "unit": "maps/s", "cores": cores, "kind": "port",
                                "sample": sample + "; oracle/db_post.py = upstream DBPostProcess restated (cv2 + restated Clipper)"}
    return line


def bench_ctc(cx: Ctx, top: bool):
    """config 4: CTC greedy decode, 1024 x 40 x 6625 per GPU per step."""
    torch = cx.torch
    args = cx.args
    from ocr_system_b200 import ops
    from ocr_system_b200.paddle_ops import CTCLabelDecode

    N, P = args.ctc_crops, max(2, args.pool // 2)
    K, W = args.steps, args.warmup
    dev, rank, world = cx.dev, cx.rank, cx.world
    pool = [ops.synth_ctc(N, CTC_T, CTC_C, crop0=(rank * P + p) * N, seed=1, device=dev) for p in range(P)]
    for i in range(W):
        ops.ctc_greedy(pool[i % P])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ops.ctc_greedy(pool[0])
    torch.cuda.synchronize()
    est = (time.perf_counter() - t0) * 1e3
    st = {"i": 0}

    def one_round():
        for _ in range(K):
            out = ops.ctc_greedy(pool[st["i"] % P])
            st["i"] += 1
            del out

    sampler = ClockSampler(_physical_index(cx.local))
    sampler.start()
    l0 = ops.launch_count()
    ms, R = cx.timed_rounds(one_round, K, est)
    launches = ops.launch_count() - l0
    clocks = sampler.stop()
    host = torch.empty(pool[0].shape, dtype=torch.float32, pin_memory=True)
    host.copy_(pool[0])
    torch.cuda.synchronize()
    dec = CTCLabelDecode(character=_ctc_chars()[1:])
    for _ in range(2):
        out = dec(host)
    Ke = max(3, min(K, 20))
    cx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(Ke):
        out = dec(host)
    e1.record()
    cx.barrier()
    e2e_ms = cx.max_over_ranks(e0.elapsed_time(e1))
    del host
    if rank != 0:
        return None
    peak, peak_src = _peaks()
    steps_timed = R * K
    alg = N * CTC_T * CTC_C * 4
    achieved = alg / (ms / steps_timed * 1e-3) / 1e9
    traffic, traffic_src = _ncu_traffic(f"ctc@{N}x{CTC_T}x{CTC_C}")
    line = {
        "metric": METRIC, "value": world * N * steps_timed / (ms / 1e3), "unit": "crops/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / steps_timed, "rounds": R, "timed_steps": steps_timed, "timed_region_s": ms / 1e3,
        "distinct_batches": P, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": _workload4(N), "decoded_first": [out[0][0][:12], out[0][1]],
        "e2e": {"value": world * N * Ke / (e2e_ms / 1e3), "unit": "crops/s", "h2d_bytes_per_step": alg,
                "d2h_bytes_per_step": N * CTC_T * 4 + N * 8, "ms_per_step": e2e_ms / Ke, "timed_steps": Ke,
                "note": "paddle_ops.CTCLabelDecode(pinned host preds [N,T,C]) -> [(utf-8 text, conf)]: H2D of the posteriors, "
                        "argmax + collapse kernels, D2H of indices / lengths / confidences, string join on the host"},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"kernel": "ctc_argmax_kernel (+ ctc_collapse_kernel)", "bound": "hbm", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": alg},
    }
    if world == 1 and not args.no_cpu_baseline:
        rate, dt, cores, sample = cpu_ctc_rate()
        line["cpu_baseline"] = {"value": rate, "unit": "crops/s", "cores": cores, "kind": "port", "sample": sample}
    return line


def bench_stream(cx: Ctx, top: bool):
    """config 5: n-page stream sharded over the ranks: chain + DB + CTC per 64-page batch, batches software-pipelined."""
    import numpy as np

    torch = cx.torch
    args = cx.args
    from ocr_system_b200 import ops
    from ocr_system_b200.pipeline import PagePipeline, shard_range

    dev, rank, world = cx.dev, cx.rank, cx.world
    n_pages, B, cpp = args.stream_pages, args.batch, args.crops_per_page
    lo, hi = shard_range(n_pages, rank, world)
    pipe = PagePipeline(max_dimension=args.max_dim, device=dev)
    tw, th = ops.target_size(PAGE_W, PAGE_H, args.max_dim)
    dh, dw = ops.det_target_size(th, tw, 960)
    torch.cuda.empty_cache()
    # two distinct resident batches (rasters + detector maps + recogniser posteriors: 36 GB each) rotated through the stream:
    # the generators stand in for the rasteriser and the two networks and stay outside the timed region
    POOL = 2
    pool = []
    for k in range(POOL):
        p0 = lo + k * B
        pool.append((ops.synth_pages(B, PAGE_H, PAGE_W, seed0=p0, device=dev), ops.synth_prob_maps(B, dh, dw, seed0=p0, device=dev),
                     ops.synth_ctc(B * cpp, CTC_T, CTC_C, crop0=p0 * cpp, seed=1, device=dev)))
    sizes = [min(B, hi - q) for q in range(lo, hi, B)]
    main = torch.cuda.current_stream(dev)
    aux = torch.cuda.Stream(dev)            # DB + CTC of batch i run beside the chain of batch i + 1

    def post_chain(i, nb, acc):
        """DBPostProcess + CTC decode of batch i on the side stream, after its chain (enqueued on `main`)."""
        ev = torch.cuda.Event()
        ev.record(main)
        aux.wait_event(ev)
        _pg, prob, post = pool[i % POOL]
        with torch.cuda.stream(aux):
            src = np.tile(np.array([[th, tw]], np.int32), (nb, 1))
            _boxes, _scores, counts = ops.db_postprocess(prob[:nb], src, DB_KW["thresh"], DB_KW["box_thresh"],
                                                         DB_KW["unclip_ratio"], DB_KW["max_candidates"], 3)
            _idx, _pos, ln, conf = ops.ctc_greedy(post[:nb * cpp])
            acc[0] += counts.sum()
            acc[1] += ln.sum()
        return counts, ln, conf

    def run_resident():
        acc = torch.zeros(2, dtype=torch.int64, device=dev)
        angle_sum = 0.0
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(main)
        for i, res in enumerate(pipe.run_device_stream(pool[j % POOL][0][:nb] for j, nb in enumerate(sizes))):
            post_chain(i, sizes[i], acc)
            angle_sum += float(np.abs(res.angles).sum())
        main.wait_stream(aux)
        b.record(main)
        torch.cuda.synchronize()
        return a.elapsed_time(b), acc.cpu().numpy(), angle_sum

    def run_from_host(host_pool, pinned):
        """the same stream with the rasters starting in pinned host memory: upload of batch i + 1 beside the kernels of
        batch i (run_host_stream), rasters + masks + angles, box counts and decoded lengths / confidences back on the host"""
        acc = torch.zeros(2, dtype=torch.int64, device=dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(main)
        for i, (_out_host, _res, _h2d, _d2h) in enumerate(pipe.run_host_stream(host_pool[j % POOL][:nb] for j, nb in enumerate(sizes))):
            nb = sizes[i]
            counts, ln, conf = post_chain(i, nb, acc)
            with torch.cuda.stream(aux):
                slot = pinned[i % 2]
                slot[0][:nb].copy_(counts, non_blocking=True)
                slot[1][:nb * cpp].copy_(ln, non_blocking=True)
                slot[2][:nb * cpp].copy_(conf, non_blocking=True)
        main.wait_stream(aux)
        b.record(main)
        torch.cuda.synchronize()
        return a.elapsed_time(b)

    # warm-up: a short stream through the same code path (plans, workspaces, stream state)
    full = sizes
    sizes = full[:min(len(full), max(2, min(args.warmup, 3)))]
    if sizes:
        run_resident()
    sizes = full
    cx.barrier()
    sampler = ClockSampler(_physical_index(cx.local))
    sampler.start()
    l0 = ops.launch_count()
    wall0 = time.perf_counter()
    hot_ms, totals, angle_sum = run_resident() if sizes else (0.0, np.zeros(2), 0.0)
    launches = ops.launch_count() - l0
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    e2e_ms = 0.0
    if args.stream_e2e:
        host_pool = pinned = None
        if sizes:
            host_pool = [torch.empty(pool[k][0].shape, dtype=torch.uint8, pin_memory=True) for k in range(POOL)]
            for k in range(POOL):
                host_pool[k].copy_(pool[k][0])
            pinned = [(torch.empty(B, dtype=torch.int32, pin_memory=True), torch.empty(B * cpp, dtype=torch.int32, pin_memory=True),
                       torch.empty(B * cpp, dtype=torch.float32, pin_memory=True)) for _ in range(2)]
            torch.cuda.synchronize()
            sizes = full[:2]
            run_from_host(host_pool, pinned)       # warm-up (pinned result slots, upload double buffer)
            sizes = full
        cx.barrier()                               # every rank, also one whose shard is empty
        if sizes:
            e2e_ms = run_from_host(host_pool, pinned)
    cx.barrier()
    hot_max = cx.max_over_ranks(hot_ms)
    e2e_max = cx.max_over_ranks(e2e_ms)
    wall_max = cx.max_over_ranks(wall)
    tot_boxes = cx.sum_over_ranks(float(totals[0]))
    tot_chars = cx.sum_over_ranks(float(totals[1]))
    if rank != 0:
        return None
    nbatches = len(full)
    line = {
        "metric": METRIC, "value": n_pages / (hot_max / 1e3), "unit": UNIT, "n_gpus": world, "steps": nbatches, "warmup": args.warmup,
        "ms_per_step": hot_max / max(nbatches, 1), "timed_region_s": hot_max / 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": _workload5(args, n_pages),
        "pages_rank0": hi - lo, "wall_s": wall_max, "boxes_total": tot_boxes, "decoded_chars_total": tot_chars,
        "mean_abs_angle_rank0": angle_sum / max(hi - lo, 1),
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if args.stream_e2e:
        line["e2e"] = {"value": n_pages / (e2e_max / 1e3), "unit": UNIT, "h2d_bytes_per_step": B * PAGE_H * PAGE_W * 3,
                       "d2h_bytes_per_step": B * th * tw * 4 + B * 8 + B * 4 + B * cpp * 8,
                       "ms_per_step": e2e_max / max(nbatches, 1),
                       "note": "the same stream with the rasters starting in pinned host memory (PagePipeline.run_host_stream: the "
                               "upload of batch i+1 beside the kernels of batch i; D2H of rasters + masks + angles), detector map / "
                               "posteriors resident (they come from the networks), DB + CTC of batch i on a side stream, D2H of box "
                               "counts and decoded lengths / confidences; whole stream from an idle pipeline"}
    return line


def main_ours(args):
    cx = Ctx(args)
    cfg = args.config
    top = None
    others = {}
    if cfg in ("2", "all"):
        top = bench_chain(cx)
    if cfg in ("3", "all"):
        r = bench_db(cx, cfg == "3")
        if cfg == "3":
            top = r
        else:
            others["3"] = r
    if cfg in ("4", "all"):
        r = bench_ctc(cx, cfg == "4")
        if cfg == "4":
            top = r
        else:
            others["4"] = r
    if cfg in ("5", "all"):
        if cfg == "all" and not args.stream_pages_set:
            args.stream_pages = min(args.stream_pages, 64 * 16 * cx.world)   # bounded in the default run
        r = bench_stream(cx, cfg == "5")
        if cfg == "5":
            top = r
        else:
            others["5"] = r
    if cx.rank == 0:
        if others:
            top["other_configs"] = others
        print(json.dumps(top), flush=True)
    if cx.world > 1:
        cx.dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="all", choices=["2", "3", "4", "5", "all"],
                    help="BASELINE.json config (1-based like the survey: 2 = configs[1]); 'all' = config 2 as the line + the others inside it")
    ap.add_argument("--batch", type=int, default=64, help="pages per GPU per step")
    ap.add_argument("--pool", type=int, default=4, help="distinct resident batches rotated through the timed steps")
    ap.add_argument("--max-dim", type=int, default=960)
    ap.add_argument("--db-maps", type=int, default=256)
    ap.add_argument("--ctc-crops", type=int, default=1024)
    ap.add_argument("--stream-pages", type=int, default=None, help="config 5: pages in the stream (default 10000)")
    ap.add_argument("--crops-per-page", type=int, default=500,
                    help="config 5: CTC crops decoded per page (config 3 density: ~500 boxes per page, one crop per box)")
    ap.add_argument("--stream-e2e", action="store_true", default=True)
    ap.add_argument("--no-stream-e2e", dest="stream_e2e", action="store_false")
    ap.add_argument("--cpu-sample", type=int, default=0, help="pages in the CPU baseline sample (default 8 x cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--traffic-bytes", type=float, default=None,
                    help="override: dram bytes/launch of the resize kernel (default: profiles/roofline_traffic.json)")
    args = ap.parse_args()
    args.stream_pages_set = args.stream_pages is not None
    if args.stream_pages is None:
        args.stream_pages = 10000
    if args.impl == "reference":
        return main_reference(args)
    args.warmup = max(args.warmup, 3)
    return main_ours(args)


if __name__ == "__main__":
    sys.exit(main())
