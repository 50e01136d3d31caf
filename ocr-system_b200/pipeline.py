"""Batched page chain + page-range sharding.

``PagePipeline`` is the throughput API behind the drop-in ``ImagePreprocessor``:
one call processes a batch of equally sized page rasters that are either already
resident in HBM (``run_device``) or in pinned host memory (``run_host``: H2D of
the rasters, kernels, D2H of the results a caller consumes).  It is the batch axis
the reference iterates one page at a time (services/ocr_service.py:604-660 under
``Semaphore(1)`` :157,404).

Stages (BASELINE.json configs[1]: gray / resize / normalize / binarize / deskew):
  a3  resize_if_needed -> PIL Lanczos to ``max_dimension``
  a10 deskew           -> cv gray + Canny + HoughLinesP + median angle + bicubic warp
  [a5+a6 contrast 1.2 + sharpness 1.1 when ``enhance=True`` (reference default chain)]
  a4  PIL gray, a9 adaptive Gaussian threshold, a15 det resize+normalize (CHW f32)

Pages are independent, so multi-GPU is plain page-range sharding with no
collective on the data path (``shard_range``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import os
import threading

import torch

from . import ops


def bind_host_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off (sysfs), so that the pinned staging
    buffers it allocates afterwards are local to that GPU's PCIe root.  With several ranks uploading 1.7 GB per
    step each, cross-socket staging memory is what limits the host side.  Returns the node, or None when the
    topology cannot be read (then nothing is changed)."""
    try:
        import os

        pr = torch.cuda.get_device_properties(device_index)
        addr = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{addr}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:  # noqa: BLE001 - a placement hint, never an error
        return None


def shard_range(n_pages: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous page range of ``rank``: [r*ceil(N/G), min(N, (r+1)*ceil(N/G)))  (SURVEY 8e)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    per = -(-n_pages // world_size)
    lo = min(n_pages, rank * per)
    return lo, min(n_pages, lo + per)


def batch_ranges(n_pages: int, rank: int, world_size: int, batch: int, resume_from: Optional[int] = None):
    """The page-index ranges ``[lo, hi)`` of ``rank``'s batches over a stream of ``n_pages`` pages: its contiguous shard
    (``shard_range``) cut into batches of ``batch`` pages, the last one ragged.  Pages are identified by their index in the
    stream, so a stream is restartable: ``resume_from`` = the first page index of this rank that is not finished yet (what a
    caller records after each yielded batch) skips the batches before it.  Recorded after a yielded batch it is one of the
    first run's batch boundaries, so the restarted run cuts the remainder into the very same batches; any other index inside
    the shard works too (pages are independent: the per-page results do not depend on the batch they travel in)."""
    if batch <= 0:
        raise ValueError("batch must be positive")
    lo, hi = shard_range(n_pages, rank, world_size)
    if resume_from is not None:
        if not (lo <= resume_from <= hi):
            raise ValueError(f"resume_from={resume_from} is outside this rank's shard [{lo}, {hi}]")
        lo = resume_from
    for q in range(lo, hi, batch):
        yield q, min(hi, q + batch)


class _StageTimer:
    """CUDA-event brackets around each stage on the launching stream (only when profiling)."""

    def __init__(self, enabled: bool):
        self.enabled = enabled
        self.marks: List[Tuple[str, torch.cuda.Event, torch.cuda.Event]] = []

    def run(self, name, fn):
        if not self.enabled:
            return fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        self.marks.append((name, a, b))
        return r

    def collect(self) -> Dict[str, float]:
        out: Dict[str, float] = {}
        for name, a, b in self.marks:
            b.synchronize()
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out


@dataclass
class PageBatchResult:
    pages: torch.Tensor          # [N,h,w,3] u8 deskewed (and enhanced) rasters
    angles: np.ndarray           # [N] float64 deskew angles (reference semantics)
    gray: torch.Tensor           # [N,h,w] u8 PIL-L of `pages`
    binary: torch.Tensor         # [N,h,w] u8 {0,255} adaptive threshold
    det_input: torch.Tensor      # [N,3,oh,ow] f32 normalised CHW detector input
    shape_list: np.ndarray       # [N,4] (src_h, src_w, ratio_h, ratio_w)
    timer: Optional["_StageTimer"] = None

    @property
    def stage_ms(self) -> Dict[str, float]:
        """Per-stage device milliseconds (CUDA events on the launch stream); synchronises."""
        return self.timer.collect() if self.timer is not None else {}


class PagePipeline:
    def __init__(self, max_dimension: int = 960, deskew: bool = True, enhance: bool = False,
                 det_limit_side_len: int = 960, device: Optional[torch.device] = None, deskew_mode: str = "exact",
                 cv_dispatch: Optional[str] = None):
        """``deskew_mode``: "exact" (default) reproduces the reference's angle bit for bit (Canny + HoughLinesP +
        median); "fast" takes it from the projection-profile estimator (``ops.estimate_skew_fast``): a different,
        tolerance-certified estimate -- rasters then differ from the reference's (tests/test_gpu_fast_skew.py)."""
        if deskew_mode not in ("exact", "fast"):
            raise ValueError("deskew_mode must be 'exact' or 'fast'")
        self.deskew_mode = deskew_mode
        self.cv_dispatch = cv_dispatch      # OpenCV build the adaptive threshold reproduces (ops.default_cv_dispatch when None)
        self.max_dimension = int(max_dimension)
        self.deskew = deskew
        self.enhance = enhance
        self.det_limit = int(det_limit_side_len)
        self.device = torch.device(device) if device is not None else None
        # streams, double buffers and pinned staging are per (pipeline object, calling thread): the reference's
        # singleton is called from asyncio.to_thread workers (ocr_service.py:674-676), so two threads may be inside
        # the same PagePipeline at once and must never share a staging buffer
        self._tls = threading.local()

    def _dev(self) -> torch.device:
        return self.device or torch.device("cuda", torch.cuda.current_device())

    # ------------------------------------------------------------------ resident input
    def run_device(self, pages: torch.Tensor, profile: bool = False) -> PageBatchResult:
        """pages: CUDA uint8 [N,H,W,3], resident in HBM."""
        if not pages.is_cuda:
            raise TypeError("run_device needs a CUDA tensor (use run_host for host buffers)")
        t = _StageTimer(profile)
        return self._back(*self._front(pages, t), t)

    def _front(self, pages: torch.Tensor, t: "_StageTimer"):
        """Everything up to the line lists (no host synchronisation)."""
        x = t.run("resize_lanczos", lambda: ops.resize_if_needed(pages, self.max_dimension))
        lines = nlines = None
        if self.deskew:
            edges = t.run("canny", lambda: ops.canny(x, 50, 150))
            if self.deskew_mode == "fast":
                lines = t.run("skew_fast", lambda: ops.estimate_skew_fast(edges))   # [N] f64 angles instead of line lists
            else:
                lines, nlines = t.run("ppht", lambda: ops.hough_lines_p(edges))
        return x, lines, nlines

    def _back(self, x, lines, nlines, t: "_StageTimer") -> PageBatchResult:
        """The reference's median / gating (the chain's one host synchronisation: the per-line angles are numpy's, like
        the reference's) and everything after it."""
        angles = np.zeros(x.shape[0], np.float64)
        if self.deskew and self.deskew_mode == "fast":
            x, angles = t.run("angle+warp", lambda: self._rotate_fast(x, lines))
        elif self.deskew:
            x, angles = t.run("angle+warp", lambda: self._rotate(x, lines, nlines))
        tail = self._tail(x, t)
        return PageBatchResult(tail[0], angles, *tail[1:], t)

    def _tail(self, x, t: "_StageTimer"):
        if self.enhance:
            x = t.run("contrast+sharpness", lambda: ops.contrast_sharpness(x, 1.2, 1.1))
        gray = t.run("gray_pil", lambda: ops.gray_pil(x))
        binary = t.run("adaptive_binarize", lambda: ops.adaptive_binarize(gray, 2, self.cv_dispatch))
        det, shape_list = t.run("det_resize_normalize", lambda: ops.det_resize_normalize(x, self.det_limit))
        return x, gray, binary, det, shape_list

    def run_device_stream(self, batches, profile: bool = False):
        """``run_device`` over a sequence of resident batches, software-pipelined on two CUDA streams: the front
        half of batch i+1 (resize, Canny, HoughLinesP) is enqueued before the host waits for the line lists of
        batch i, so its kernels fill the SMs that batch i's last HoughLinesP clusters leave idle (49 pages fit at
        once; a 64-page batch alone runs 1.3 waves).  Yields one PageBatchResult per batch, in order; the
        results are safe to use on the caller's current stream."""
        dev = self._dev()
        main = torch.cuda.current_stream(dev)
        st = getattr(self._tls, "dev_streams", None)
        if st is None or st[0].device != dev:
            # one high-priority stream for the wide kernels (resize, Canny, warp, gray, binarize, normalise) and two
            # low-priority streams that alternate for HoughLinesP: when a page's cluster retires, the freed SMs go to
            # the wide kernels of the next batch first, so its HoughLinesP is ready to follow without a gap
            st = self._tls.dev_streams = (torch.cuda.Stream(dev, priority=-1), torch.cuda.Stream(dev, priority=0),
                                          torch.cuda.Stream(dev, priority=0))
        hi = st[0]

        def start(i, pages):
            lo = st[1 + (i & 1)]
            t = _StageTimer(profile)
            hi.wait_stream(main)
            pages.record_stream(hi)               # the caller may drop the batch while hi still reads it
            with torch.cuda.stream(hi):
                x = t.run("resize_lanczos", lambda: ops.resize_if_needed(pages, self.max_dimension))
                job = None
                lines = nlines = None
                if self.deskew:
                    edges = t.run("canny", lambda: ops.canny(x, 50, 150))
                    if self.deskew_mode == "fast":
                        lines = t.run("skew_fast", lambda: ops.estimate_skew_fast(edges))
                    else:
                        job = t.run("ppht_prepare", lambda: ops.HoughJob(edges).prepare())
            if job is not None:
                lo.wait_stream(hi)
                with torch.cuda.stream(lo):
                    lines, nlines = t.run("ppht", job.lines)
            return lo, (x, lines, nlines), t

        def finish(job):
            lo, front, t = job
            hi.wait_stream(lo)                    # the line lists of this batch; the host waits for them inside _back
            with torch.cuda.stream(hi):
                res = self._back(*front, t)
            main.wait_stream(hi)
            for ten in (res.pages, res.gray, res.binary, res.det_input):
                ten.record_stream(main)
            return res

        depth = max(1, int(os.environ.get("LUMINA_STREAM_DEPTH", "2")))   # batches enqueued ahead of the one being finished
        pending = []
        for i, pages in enumerate(batches):
            if not pages.is_cuda:
                raise TypeError("run_device_stream needs CUDA tensors")
            pending.append(start(i, pages))
            if len(pending) > depth - 1 and len(pending) > 1:
                yield finish(pending.pop(0))
        while pending:
            yield finish(pending.pop(0))

    def _rotate_fast(self, x: torch.Tensor, est: torch.Tensor):
        """deskew_mode="fast": the reference's gates (image_preprocessing.py:433-439) and warp on the estimated angles."""
        n, h, w = x.shape[0], x.shape[1], x.shape[2]
        a = est.cpu().numpy()     # n doubles; synchronises like the exact path's line lists
        angles = np.where(np.abs(a) > 45, 0.0, a)
        apply = ((np.abs(a) >= 0.5) & (np.abs(a) <= 45)).astype(np.uint8)
        mats = np.zeros((n, 6), np.float64)
        for i in np.nonzero(apply)[0]:
            mats[i] = ops.rotation_matrix(w // 2, h // 2, float(a[i]), 1.0).reshape(-1)
        if apply.any():
            x = ops.warp_affine_cubic(x, mats, apply)
        return x, angles

    def _rotate(self, x: torch.Tensor, lines: torch.Tensor, nlines: torch.Tensor):
        """Host median / gating (image_preprocessing.py:414-439) + one warp launch.  The only
        host synchronisation of the chain: the line lists (a few KB per page) come back."""
        n, h, w = x.shape[0], x.shape[1], x.shape[2]
        # pinned staging; the counts come back first (256 bytes), then only the line slots in use (a text page has a
        # few hundred segments of the 4096 slots: ~0.4 MB per 64 pages instead of 4 MB)
        key = (n, lines.shape[1], x.device)
        st = getattr(self._tls, "rotate_stage", None)   # per thread, shape-keyed
        if st is None or st[0] != key:
            st = self._tls.rotate_stage = (key, torch.empty(n, dtype=torch.int32, pin_memory=True),
                                           torch.empty((n, lines.shape[1], 4), dtype=torch.int32, pin_memory=True))
        _, nl_pin, ln_pin = st
        cur = torch.cuda.current_stream(x.device)
        nl_pin.copy_(nlines, non_blocking=True)
        cur.synchronize()
        nl = nl_pin.numpy()
        keep = int(nl.max(initial=0))
        if keep > lines.shape[1]:  # truncated list: redo the Hough stage with room for every line
            edges = ops.canny(x, 50, 150)
            lines, nlines = ops.hough_lines_p(edges, max_lines=keep)
            nl = nlines.cpu().numpy()
            lh = lines.cpu().numpy()
        elif keep == 0:
            lh = ln_pin.numpy()[:, :1]
        else:
            ops.copy_lines_to_host(lines, keep, ln_pin)       # [n][keep][4] packed at the front of the pinned buffer
            cur.synchronize()
            lh = ln_pin.numpy().reshape(-1)[: n * keep * 4].reshape(n, keep, 4)
        angles, mats, apply = ops.deskew_decide(lh, nl, h, w)
        if apply.any():
            x = ops.warp_affine_cubic(x, mats, apply)
        return x, angles

    # ------------------------------------------------------------------ encoded input (files cross PCIe, not rasters)
    def run_encoded_stream(self, encoded_batches, profile: bool = False, keep: int = 2):
        """The e2e path for *files*: ``encoded_batches`` yields ``(blob, offsets)`` -- a (pinned) CPU uint8 tensor
        holding the baseline-JPEG files of one batch back to back and the int64 offsets [n+1] -- i.e. what
        ``load_image_bytes`` (image_preprocessing.py:70-75) is handed, batched.  The files are uploaded (1-2 MB
        per A4 page instead of the 26 MB raster) and decoded in HBM on a side stream (``ops.JpegDecoder``: the
        raster equals Pillow's), then go through the stream-pipelined chain (``run_device_stream``); the results a
        caller consumes (deskewed rasters, masks, angles) are copied back to pinned host memory.  Yields
        ``(out_host, PageBatchResult, h2d_bytes, d2h_bytes)`` per batch, complete in host memory; ``out_host``
        rotates over ``keep`` slots as in ``run_host_stream``.  A page whose entropy-coded data is corrupt raises
        ``LuminaError`` (decode such a file on the host, as the reference does)."""
        dev = self._dev()
        keep = max(1, int(keep))
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            st = getattr(self._tls, "enc_state", None)
            if st is None or st["dev"] != dev:
                st = self._tls.enc_state = {"dev": dev, "stream": torch.cuda.Stream(dev, priority=-1),
                                            "dec": ops.JpegDecoder(), "out": []}
            dec_stream, dec = st["stream"], st["dec"]
            meta = []

            def decoded():
                for blob, offs in encoded_batches:
                    with torch.cuda.stream(dec_stream):
                        pages, status = dec.decode(blob, offs, dev)
                        ev = torch.cuda.Event()
                        ev.record(dec_stream)
                    main.wait_event(ev)            # run_device_stream's streams wait for `main`
                    pages.record_stream(main)
                    status.record_stream(main)
                    meta.append((status, int(offs[-1] - offs[0])))
                    yield pages

            outs = st["out"]
            for i, res in enumerate(self.run_device_stream(decoded(), profile=profile)):
                status, nbytes = meta.pop(0)
                if len(outs) != keep or outs[0]["pages"].shape != res.pages.shape or outs[0]["binary"].shape != res.binary.shape:
                    outs[:] = [dict(self._new_out(res), status=torch.empty(status.shape, dtype=torch.int32, pin_memory=True))
                               for _ in range(keep)]
                out_host = outs[i % keep]
                out_host["pages"].copy_(res.pages, non_blocking=True)
                out_host["binary"].copy_(res.binary, non_blocking=True)
                out_host["status"].copy_(status, non_blocking=True)
                main.synchronize()
                bad = np.nonzero(out_host["status"].numpy())[0]
                if bad.size:
                    raise ops._abi.LuminaError(f"corrupt entropy-coded data in pages {bad.tolist()} of batch {i}")
                out_host["angles"] = res.angles
                yield out_host, res, nbytes, res.pages.numel() + res.binary.numel() + res.angles.nbytes + 4 * status.numel()

    # ------------------------------------------------------------------ host input (the e2e path)
    @staticmethod
    def _new_out(res: PageBatchResult) -> Dict[str, torch.Tensor]:
        return {"pages": torch.empty(res.pages.shape, dtype=torch.uint8, pin_memory=True),
                "binary": torch.empty(res.binary.shape, dtype=torch.uint8, pin_memory=True)}

    def run_host(self, pages_host: torch.Tensor, out_host: Optional[Dict[str, torch.Tensor]] = None,
                 profile: bool = False):
        """pages_host: (pinned) CPU uint8 [N,H,W,3].  Copies the rasters to the device, runs the
        chain and copies back what a caller consumes: deskewed rasters, binary masks and angles.
        Returns (out_host dict, PageBatchResult, h2d_bytes, d2h_bytes); ``out_host`` is freshly allocated
        (owned by the caller) unless one is passed in."""
        if pages_host.is_cuda:
            raise TypeError("run_host needs a host tensor")
        dev = self._dev()
        with torch.cuda.device(dev):
            x = pages_host.to(dev, non_blocking=True)
            res = self.run_device(x, profile=profile)
            if out_host is None:
                out_host = self._new_out(res)
            out_host["pages"].copy_(res.pages, non_blocking=True)
            out_host["binary"].copy_(res.binary, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
        out_host["angles"] = res.angles
        h2d = pages_host.numel()
        d2h = res.pages.numel() + res.binary.numel() + res.angles.nbytes
        return out_host, res, h2d, d2h

    def run_host_stream(self, host_batches, profile: bool = False, keep: int = 2):
        """Streaming form of ``run_host`` for a sequence of (pinned) host batches: the H2D copy of
        batch i+1 is enqueued on a copy stream before batch i is processed, so raster upload overlaps
        the kernels of the previous batch (double-buffered device input).  Yields
        ``(out_host, PageBatchResult, h2d_bytes, d2h_bytes)`` per batch, results complete in host memory.

        Lifetime of a yielded result: the pinned ``out_host`` buffers rotate over ``keep`` slots, so batch i's
        host results stay intact until the generator has been advanced ``keep`` more times (``keep=2``: a consumer
        may still hold batch i while it receives batch i+1).  The device tensors of ``PageBatchResult`` are fresh
        allocations and follow the usual torch lifetime; the rasters never alias the reused upload slot.  Copy what
        must live longer (or raise ``keep``)."""
        dev = self._dev()
        keep = max(1, int(keep))
        with torch.cuda.device(dev):
            comp = torch.cuda.current_stream(dev)
            st = getattr(self._tls, "stream_state", None)
            if st is None or st["dev"] != dev:  # copy stream + double buffer live as long as the (pipeline, thread) pair
                st = self._tls.stream_state = {"dev": dev, "copy": torch.cuda.Stream(dev), "buf": [None, None],
                                               "free": [None, None], "out": []}
            copy_stream = st["copy"]
            it = iter(host_batches)
            dev_buf = st["buf"]         # persistent double buffer for the rasters (no allocator traffic per batch)
            ready = [None, None]        # H2D of slot k finished (recorded on the copy stream)
            free = st["free"]           # kernels that read slot k finished (recorded on the compute stream)
            host_of = [None, None]

            def issue(k, hb):
                if hb.is_cuda:
                    raise TypeError("run_host_stream needs host tensors")
                if dev_buf[k] is None or dev_buf[k].shape != hb.shape:
                    dev_buf[k] = torch.empty(hb.shape, dtype=torch.uint8, device=dev)
                    free[k] = None
                    # a fresh slot comes from the compute stream's allocator pool: the block may belong to a tensor
                    # whose last kernels are still queued there, so the first upload into it is ordered behind them
                    copy_stream.wait_stream(comp)
                with torch.cuda.stream(copy_stream):
                    if free[k] is not None:
                        copy_stream.wait_event(free[k])
                    dev_buf[k].copy_(hb, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                ready[k] = ev
                host_of[k] = hb

            cur_hb = next(it, None)
            if cur_hb is None:
                return
            issue(0, cur_hb)
            i = 0
            outs = st["out"]  # pinned result slots are kept too: cudaHostAlloc costs ~100 ms per call otherwise
            while cur_hb is not None:
                k = i & 1
                nxt = next(it, None)
                if nxt is not None:
                    issue(k ^ 1, nxt)  # upload of the next batch overlaps this batch's kernels
                comp.wait_event(ready[k])
                res = self.run_device(dev_buf[k], profile=profile)
                if res.pages.data_ptr() == dev_buf[k].data_ptr():
                    res.pages = res.pages.clone()   # nothing resized or rotated: do not hand out the reused upload slot
                fe = torch.cuda.Event()
                fe.record(comp)
                free[k] = fe
                if len(outs) != keep or outs[0]["pages"].shape != res.pages.shape or outs[0]["binary"].shape != res.binary.shape:
                    outs[:] = [self._new_out(res) for _ in range(keep)]
                out_host = outs[i % keep]
                out_host["pages"].copy_(res.pages, non_blocking=True)
                out_host["binary"].copy_(res.binary, non_blocking=True)
                comp.synchronize()
                out_host["angles"] = res.angles
                hb = host_of[k]
                yield out_host, res, hb.numel(), res.pages.numel() + res.binary.numel() + res.angles.nbytes
                cur_hb = nxt
                i += 1
