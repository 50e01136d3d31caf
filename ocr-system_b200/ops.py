"""Batched page-tensor operators over the C-ABI (``include/lumina_b200.h``).

Every function takes CUDA ``torch.uint8`` page batches ``[N, H, W, C]`` (C in
{1, 3}; planes are ``[N, H, W]``), allocates outputs / scratch as torch tensors
(torch is only the allocator + stream provider) and enqueues the hand-written
sm_100a kernels on the current torch stream.  No CPU fallback exists: a
non-CUDA tensor is an error.
"""
from __future__ import annotations

import collections
import contextlib
import ctypes as C
import functools
import os
import threading
from typing import Optional, Tuple

import numpy as np
import torch

from . import _abi

_L = _abi.lib
_chk = _abi.check


def _ptr(t) -> C.c_void_p:
    return C.c_void_p(t.data_ptr() if t is not None else 0)


def _stream() -> C.c_void_p:
    """The current torch stream of the CURRENT device.  Every launching function below is wrapped in
    ``_on_tensor_device``, which makes the tensor's device current first -- so this is always the stream of the
    device the pointers live on (the library's per-device tables key off cudaGetDevice the same way)."""
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _on_tensor_device(fn):
    """Run ``fn`` with the device of its first CUDA tensor argument current.  A cuda:1 tensor handed to a thread
    whose current device is cuda:0 therefore launches on cuda:1's context and stream, not on cuda:0's."""

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        for a in args:
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args, **kwargs)
                break
        return fn(*args, **kwargs)

    return wrapper


def _pages(x: torch.Tensor) -> Tuple[torch.Tensor, int, int, int, int]:
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise TypeError("lumina_b200 ops need CUDA tensors (there is no CPU fallback)")
    if x.dtype != torch.uint8:
        raise TypeError(f"expected uint8 pages, got {x.dtype}")
    if x.dim() == 3:
        x = x.unsqueeze(-1)
    if x.dim() != 4 or x.shape[-1] not in (1, 3):
        raise ValueError(f"expected [N,H,W,C] with C in (1,3) or [N,H,W]; got {tuple(x.shape)}")
    x = x.contiguous()
    if x.data_ptr() & 15:      # a slice such as pages[1:] of odd-sized pages: the 128-bit kernels need 16-byte bases
        x = x.clone()
    n, h, w, c = x.shape
    return x, n, h, w, c


def _like(x: torch.Tensor, squeeze: bool) -> torch.Tensor:
    return x.squeeze(-1) if squeeze else x


def launch_count() -> int:
    return int(_L().lumina_launch_count())


# --------------------------------------------------------------------------- a2
@_on_tensor_device
def exif_transpose(pages: torch.Tensor, orientation: int) -> torch.Tensor:
    """PIL ImageOps.exif_transpose (image_preprocessing.py:173) for a fixed EXIF orientation."""
    sq = pages.dim() == 3
    x, n, h, w, c = _pages(pages)
    oh, ow = (w, h) if 5 <= orientation <= 8 else (h, w)
    out = torch.empty((n, oh, ow, c), dtype=torch.uint8, device=x.device)
    _chk(_L().lumina_exif_transpose_u8(_ptr(x), _ptr(out), n, h, w, c, int(orientation), _stream()))
    return _like(out, sq)


# --------------------------------------------------------------------------- a3
def target_size(width: int, height: int, max_dim: int) -> Tuple[int, int]:
    ow, oh = C.c_int(), C.c_int()
    _L().lumina_target_size(width, height, max_dim, C.byref(ow), C.byref(oh))
    return ow.value, oh.value


class _ResizePlans:
    """Device coefficient tables per (device, in, out) geometry, created on first use and kept in a bounded LRU: a
    long-lived service sees arbitrary scan sizes, and every geometry holds a few hundred KB of device tables.  A plan
    is only evicted while no caller is between ``use().__enter__`` and the end of its launch call; kernels that were
    launched with it and are still running are covered by ``lumina_resize_plan_destroy`` itself (its ``cudaFree``s wait
    for the device).  ``LUMINA_RESIZE_PLAN_CACHE`` sets the bound (default 256 geometries)."""

    def __init__(self, capacity: Optional[int] = None, create=None, destroy=None):
        self._plans = collections.OrderedDict()      # key -> [handle, callers inside use()]
        self._lock = threading.Lock()
        self.capacity = max(1, int(capacity if capacity is not None else os.environ.get("LUMINA_RESIZE_PLAN_CACHE", "256")))
        self._create = create or self._create_native
        self._destroy = destroy or self._destroy_native

    @staticmethod
    def _create_native(dev, in_h, in_w, out_h, out_w):
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _chk(_L().lumina_resize_plan_create(in_h, in_w, out_h, out_w, C.byref(h)))
        return h

    @staticmethod
    def _destroy_native(dev, handle):
        with torch.cuda.device(dev):
            _L().lumina_resize_plan_destroy(handle)

    @contextlib.contextmanager
    def use(self, dev: int, in_h: int, in_w: int, out_h: int, out_w: int):
        key = (dev, in_h, in_w, out_h, out_w)
        with self._lock:
            ent = self._plans.get(key)
            if ent is None:
                ent = self._plans[key] = [self._create(*key), 0]
            self._plans.move_to_end(key)
            ent[1] += 1
            victims = []
            if len(self._plans) > self.capacity:
                for k in list(self._plans):
                    if len(self._plans) - len(victims) <= self.capacity:
                        break
                    if self._plans[k][1] == 0:
                        victims.append(k)
                victims = [(k, self._plans.pop(k)[0]) for k in victims]
        try:
            for k, h in victims:       # outside the lock: destroy waits for the device
                self._destroy(k[0], h)
            yield ent[0]
        finally:
            with self._lock:
                ent[1] -= 1

    def __len__(self) -> int:
        return len(self._plans)


_plans = _ResizePlans()


@_on_tensor_device
def resize_lanczos(pages: torch.Tensor, out_w: int, out_h: int) -> torch.Tensor:
    """PIL Image.resize((out_w,out_h), LANCZOS) (image_preprocessing.py:110), byte-exact."""
    sq = pages.dim() == 3
    x, n, h, w, c = _pages(pages)
    out = torch.empty((n, out_h, out_w, c), dtype=torch.uint8, device=x.device)
    with _plans.use(x.device.index, h, w, out_h, out_w) as plan:
        wsb = int(_L().lumina_resize_workspace_bytes(plan, n, c))
        ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=x.device) if wsb else None
        _chk(_L().lumina_resize_lanczos_u8(plan, _ptr(x), _ptr(out), n, c, _ptr(ws), wsb, _stream()))
    return _like(out, sq)


@_on_tensor_device
def resize_nearest(planes: torch.Tensor, out_w: int, out_h: int) -> torch.Tensor:
    """Pillow's resize of mode "P" / "1" images (always NEAREST): planes [N,H,W] of single-byte pixels."""
    x, n, h, w, c = _pages(planes)
    if c != 1:
        raise ValueError("resize_nearest works on single-byte planes [N,H,W]")
    xt = np.empty(out_w, np.int32)
    yt = np.empty(out_h, np.int32)
    _L().lumina_nearest_table_host(w, out_w, xt.ctypes.data_as(C.c_void_p))
    _L().lumina_nearest_table_host(h, out_h, yt.ctypes.data_as(C.c_void_p))
    dxt, dyt = torch.from_numpy(xt).to(x.device), torch.from_numpy(yt).to(x.device)
    out = torch.empty((n, out_h, out_w), dtype=torch.uint8, device=x.device)
    _chk(_L().lumina_resize_nearest_u8(_ptr(x), _ptr(out), n, h, w, out_h, out_w, _ptr(dxt), _ptr(dyt), _stream()))
    return out


def resize_if_needed(pages: torch.Tensor, max_dim: int) -> torch.Tensor:
    h, w = pages.shape[1], pages.shape[2]
    if max(w, h) <= max_dim:
        return pages
    ow, oh = target_size(w, h, max_dim)
    if ow < 1 or oh < 1:      # a strip so thin that its short side rounds to 0: Pillow's Image.resize raises this
        raise ValueError("height and width must be > 0")
    return resize_lanczos(pages, ow, oh)


# --------------------------------------------------------------------------- a4
@_on_tensor_device
def gray_pil(pages: torch.Tensor) -> torch.Tensor:
    x, n, h, w, c = _pages(pages)
    if c == 1:
        return x.squeeze(-1)
    out = torch.empty((n, h, w), dtype=torch.uint8, device=x.device)
    _chk(_L().lumina_rgb2gray_pil_u8(_ptr(x), _ptr(out), n * h * w, _stream()))
    return out


@_on_tensor_device
def gray_cv(pages: torch.Tensor) -> torch.Tensor:
    x, n, h, w, c = _pages(pages)
    if c == 1:
        return x.squeeze(-1)
    out = torch.empty((n, h, w), dtype=torch.uint8, device=x.device)
    _chk(_L().lumina_rgb2gray_cv_u8(_ptr(x), _ptr(out), n * h * w, _stream()))
    return out


# --------------------------------------------------------------------------- a5 / a6
@_on_tensor_device
def contrast_mean(pages: torch.Tensor) -> torch.Tensor:
    x, n, h, w, c = _pages(pages)
    scratch = torch.empty(n, dtype=torch.int64, device=x.device)
    mean = torch.empty(n, dtype=torch.int32, device=x.device)
    _chk(_L().lumina_contrast_mean_u8(_ptr(x), n, h, w, c, _ptr(scratch), _ptr(mean), _stream()))
    return mean


@_on_tensor_device
def enhance_contrast(pages: torch.Tensor, factor: float, mean: Optional[torch.Tensor] = None) -> torch.Tensor:
    sq = pages.dim() == 3
    x, n, h, w, c = _pages(pages)
    if mean is None:
        mean = contrast_mean(x)
    out = torch.empty_like(x)
    _chk(_L().lumina_contrast_apply_u8(_ptr(x), _ptr(out), n, h, w, c, _ptr(mean), float(factor), _stream()))
    return _like(out, sq)


@_on_tensor_device
def enhance_sharpness(pages: torch.Tensor, factor: float) -> torch.Tensor:
    sq = pages.dim() == 3
    x, n, h, w, c = _pages(pages)
    out = torch.empty_like(x)
    _chk(_L().lumina_sharpness_u8(_ptr(x), _ptr(out), n, h, w, c, float(factor), _stream()))
    return _like(out, sq)


@_on_tensor_device
def contrast_sharpness(pages: torch.Tensor, contrast: float, sharpness: float) -> torch.Tensor:
    """enhance_sharpness(enhance_contrast(x, contrast), sharpness) with the contrast LUT fused
    into the stencil (image_preprocessing.py:613-618)."""
    sq = pages.dim() == 3
    x, n, h, w, c = _pages(pages)
    mean = contrast_mean(x)
    out = torch.empty_like(x)
    _chk(_L().lumina_contrast_sharpness_u8(_ptr(x), _ptr(out), n, h, w, c, _ptr(mean), float(contrast),
                                           float(sharpness), _stream()))
    return _like(out, sq)


# --------------------------------------------------------------------------- a7 / a8 / a9
@_on_tensor_device
def median3(pages: torch.Tensor) -> torch.Tensor:
    sq = pages.dim() == 3
    x, n, h, w, c = _pages(pages)
    out = torch.empty_like(x)
    _chk(_L().lumina_median3_u8(_ptr(x), _ptr(out), n, h, w, c, _stream()))
    return _like(out, sq)


@_on_tensor_device
def binarize(pages: torch.Tensor, threshold: int = 128) -> torch.Tensor:
    x, n, h, w, c = _pages(pages)
    out = torch.empty((n, h, w), dtype=torch.uint8, device=x.device)
    _chk(_L().lumina_binarize_u8(_ptr(x), _ptr(out), n * h * w, c, int(threshold), _stream()))
    return out


@_on_tensor_device
def otsu_binarize(gray: torch.Tensor):
    """Per-page Otsu threshold + mask of gray planes [N,H,W] (cv2.threshold(..., THRESH_BINARY | THRESH_OTSU), bit-equal).
    Returns (mask uint8 {0,255} [N,H,W], thresholds int32 [N])."""
    if not gray.is_cuda or gray.dtype != torch.uint8 or gray.dim() != 3:
        raise TypeError("otsu_binarize expects a CUDA uint8 [N,H,W] tensor")
    g = gray.contiguous()
    n, h, w = g.shape
    out = torch.empty_like(g)
    thr = torch.empty(n, dtype=torch.int32, device=g.device)
    hist = torch.empty(n * 256, dtype=torch.int32, device=g.device)
    _chk(_L().lumina_otsu_u8(_ptr(g), _ptr(out), n, h, w, _ptr(thr), _ptr(hist), _stream()))
    return out, thr


@_on_tensor_device
def sauvola_binarize(gray: torch.Tensor, window: int = 25, k: float = 0.2, r: float = 128.0) -> torch.Tensor:
    """Sauvola local threshold of gray planes [N,H,W]: 255 where x > m * (1 + k * (s / r - 1)) over a
    window x window neighbourhood clipped to the page (integral images per tile in shared memory)."""
    if not gray.is_cuda or gray.dtype != torch.uint8 or gray.dim() != 3:
        raise TypeError("sauvola_binarize expects a CUDA uint8 [N,H,W] tensor")
    g = gray.contiguous()
    n, h, w = g.shape
    out = torch.empty_like(g)
    _chk(_L().lumina_sauvola_u8(_ptr(g), _ptr(out), n, h, w, int(window), float(k), float(r), _stream()))
    return out


CV_DISPATCH = {"plain": 0, "avx2": 1}   # LUMINA_CV_PLAIN / LUMINA_CV_AVX2 (include/lumina_b200.h)


def default_cv_dispatch() -> str:
    """Which OpenCV build the float Gaussian of adaptiveThreshold reproduces when a caller does not say: "avx2" --
    OpenCV's default dispatch on x86 hosts with AVX2 + FMA3, i.e. what the reference runs (it never calls
    cv2.setUseOptimized) -- unless LUMINA_CV_DISPATCH=plain asks for OpenCV's plain path."""
    v = os.environ.get("LUMINA_CV_DISPATCH", "avx2")
    if v not in CV_DISPATCH:
        raise ValueError(f"LUMINA_CV_DISPATCH must be one of {sorted(CV_DISPATCH)}, not {v!r}")
    return v


@_on_tensor_device
def adaptive_binarize(pages: torch.Tensor, cval: int = 2, cv_dispatch: Optional[str] = None) -> torch.Tensor:
    """cv2.adaptiveThreshold(GAUSSIAN_C, BINARY, 11, cval) of gray planes [N,H,W] (or RGB pages: PIL gray fused),
    bit-equal to OpenCV in the named dispatch mode (``default_cv_dispatch`` when None)."""
    x, n, h, w, c = _pages(pages)
    mode = CV_DISPATCH[cv_dispatch if cv_dispatch is not None else default_cv_dispatch()]
    out = torch.empty((n, h, w), dtype=torch.uint8, device=x.device)
    _chk(_L().lumina_adaptive_gauss11_ex_u8(_ptr(x), _ptr(out), n, h, w, c, int(cval), mode, _stream()))
    return out


# --------------------------------------------------------------------------- a10
def _ws(nbytes: int, device) -> torch.Tensor:
    # torch's caching allocator returns >=512-byte aligned blocks
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


@_on_tensor_device
def canny(pages: torch.Tensor, low: int = 50, high: int = 150) -> torch.Tensor:
    """cv gray (for RGB) + cv2.Canny(low, high, apertureSize=3) -> edges [N,H,W] {0,255}."""
    x, n, h, w, c = _pages(pages)
    edges = torch.empty((n, h, w), dtype=torch.uint8, device=x.device)
    wsb = int(_L().lumina_canny_workspace_bytes(n, h, w))
    ws = _ws(wsb, x.device)
    _chk(_L().lumina_canny_u8(_ptr(x), _ptr(edges), n, h, w, c, int(low), int(high), _ptr(ws), wsb, _stream()))
    return edges


@_on_tensor_device
def hough_lines_p(edges: torch.Tensor, rho: float = 1.0, theta: float = float(np.pi / 180), threshold: int = 100,
                  min_line_length: int = 100, max_line_gap: int = 10, max_lines: int = 4096):
    """cv2.HoughLinesP on a batch of edge planes -> (lines [N,max_lines,4] int32, nlines [N] int32), device."""
    x, n, h, w, c = _pages(edges)
    if c != 1:
        raise ValueError("edges must be [N,H,W]")
    lines = torch.empty((n, max_lines, 4), dtype=torch.int32, device=x.device)
    nlines = torch.empty(n, dtype=torch.int32, device=x.device)
    wsb = int(_L().lumina_ppht_workspace_bytes(n, h, w, float(rho), float(theta)))
    ws = _ws(wsb, x.device)
    _chk(_L().lumina_ppht(_ptr(x), n, h, w, float(rho), float(theta), int(threshold), int(min_line_length),
                          int(max_line_gap), _ptr(lines), _ptr(nlines), int(max_lines), _ptr(ws), wsb, _stream()))
    return lines, nlines


class HoughJob:
    """cv2.HoughLinesP in two halves for stream pipelines: ``prepare()`` (point collection, bitmask, visiting
    order: wide, short kernels) and ``lines()`` (the long cluster kernel), each on the stream that is current
    when it is called.  ``lines()`` must be ordered after ``prepare()`` by the caller (same stream or an event)."""

    def __init__(self, edges: torch.Tensor, rho: float = 1.0, theta: float = float(np.pi / 180), threshold: int = 100,
                 min_line_length: int = 100, max_line_gap: int = 10, max_lines: int = 4096):
        x, n, h, w, c = _pages(edges)
        if c != 1:
            raise ValueError("edges must be [N,H,W]")
        self.x, self.n, self.h, self.w = x, n, h, w
        self.args = (float(rho), float(theta))
        self.params = (int(threshold), int(min_line_length), int(max_line_gap))
        self.max_lines = int(max_lines)
        self.wsb = int(_L().lumina_ppht_workspace_bytes(n, h, w, *self.args))
        self.ws = _ws(self.wsb, x.device)

    def prepare(self):
        with torch.cuda.device(self.x.device):
            _chk(_L().lumina_ppht_prepare(_ptr(self.x), self.n, self.h, self.w, *self.args, _ptr(self.ws), self.wsb, _stream()))
        return self

    def lines(self):
        lines = torch.empty((self.n, self.max_lines, 4), dtype=torch.int32, device=self.x.device)
        nlines = torch.empty(self.n, dtype=torch.int32, device=self.x.device)
        with torch.cuda.device(self.x.device):
            _chk(_L().lumina_ppht_lines(_ptr(self.x), self.n, self.h, self.w, *self.args, *self.params, _ptr(lines), _ptr(nlines),
                                        self.max_lines, _ptr(self.ws), self.wsb, _stream()))
        cur = torch.cuda.current_stream(self.x.device)
        self.ws.record_stream(cur)
        self.x.record_stream(cur)
        return lines, nlines


@_on_tensor_device
def rgbx_to_rgb(pages_rgbx: torch.Tensor) -> torch.Tensor:
    """[N,H,W,4] uint8 (Pillow's in-memory R,G,B,pad) -> [N,H,W,3]."""
    if not pages_rgbx.is_cuda or pages_rgbx.dtype != torch.uint8 or pages_rgbx.dim() != 4 or pages_rgbx.shape[-1] != 4:
        raise TypeError("rgbx_to_rgb expects a CUDA uint8 [N,H,W,4] tensor")
    x = pages_rgbx.contiguous()
    out = torch.empty(x.shape[:3] + (3,), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _chk(_L().lumina_rgbx_to_rgb_u8(_ptr(x), _ptr(out), x.shape[0] * x.shape[1] * x.shape[2], _stream()))
    return out


def copy_lines_to_host(lines: torch.Tensor, keep: int, out_pinned: torch.Tensor) -> None:
    """The first ``keep`` segments of every page of a Hough result [N,max_lines,4] into the front of a pinned int32
    buffer (viewed as [N,keep,4] by the caller) -- one strided copy on the current stream of the lists' device."""
    n, stride = int(lines.shape[0]), int(lines.shape[1])
    if not lines.is_contiguous() or lines.dtype != torch.int32 or out_pinned.numel() < n * keep * 4:
        raise ValueError("copy_lines_to_host: contiguous int32 [N,max_lines,4] and a large enough pinned buffer expected")
    with torch.cuda.device(lines.device):
        _chk(_L().lumina_copy_lines_to_host(_ptr(lines), n, stride, int(keep), out_pinned.data_ptr(), _stream()))


def line_angles(lines_host: np.ndarray) -> np.ndarray:
    """image_preprocessing.py:421-426 for an array of segments [...,4]: ``np.degrees(np.arctan2(y2 - y1, x2 - x1))``
    folded to +-45 -- evaluated BY NUMPY, on int32 operands like the reference's scalars.  numpy's arctan2 is not
    glibc's on AVX-512 builds (bundled SIMD math; the last place differs for ~0.3 % of segments), and numpy's
    vectorised result equals its scalar one, so this is the reference's value on whatever host it runs."""
    ln = np.asarray(lines_host, dtype=np.int32)
    a = np.arctan2(ln[..., 3] - ln[..., 1], ln[..., 2] - ln[..., 0])
    np.degrees(a, out=a)
    lo, hi = a < -45, a > 45            # the reference's if / elif on the unfolded value; the two sets are disjoint
    np.add(a, 90, out=a, where=lo)
    np.subtract(a, 90, out=a, where=hi)
    return a


def median_angle(lines_host: np.ndarray) -> float:
    """image_preprocessing.py:414-428 on the host: per-line angles by numpy (``line_angles``), ``np.median``'s value."""
    ln = np.ascontiguousarray(lines_host, dtype=np.int32).reshape(-1, 4)
    if ln.shape[0] == 0:
        return 0.0
    return float(deskew_decide(ln[None], np.array([ln.shape[0]], np.int32), 0, 0, gate=False)[0][0])


def rotation_matrix(cx: float, cy: float, angle: float, scale: float = 1.0) -> np.ndarray:
    m = np.zeros(6, np.float64)
    _L().lumina_rotation_matrix_host(float(cx), float(cy), float(angle), float(scale), m.ctypes.data_as(C.c_void_p))
    return m.reshape(2, 3)


def deskew_decide(lines_host: np.ndarray, nlines_host: np.ndarray, h: int, w: int, gate: bool = True):
    """Reference gating (image_preprocessing.py:409-444) for a batch in one host call:
    (angles[N] f64, forward matrices[N,6] f64, apply[N] u8).  The per-line angles come from numpy (``line_angles``),
    median / gates / getRotationMatrix2D from the library.  ``gate=False`` returns the raw medians."""
    ln = np.asarray(lines_host, dtype=np.int32)
    nl = np.ascontiguousarray(nlines_host, dtype=np.int32)
    n = ln.shape[0]
    keep = max(1, min(int(nl.max(initial=0)), ln.shape[1]))
    la = np.ascontiguousarray(line_angles(ln[:, :keep]))
    angles = np.zeros(n, np.float64)
    mats = np.zeros((n, 6), np.float64)
    apply = np.zeros(n, np.uint8)
    if not gate:
        for i in range(n):
            k = min(int(nl[i]), keep)
            angles[i] = float(np.median(la[i, :k])) if k > 0 else 0.0
        return angles, mats, apply
    _L().lumina_deskew_decide_angles_host(la.ctypes.data_as(C.c_void_p), nl.ctypes.data_as(C.c_void_p), n, keep, int(h), int(w),
                                          angles.ctypes.data_as(C.c_void_p), mats.ctypes.data_as(C.c_void_p),
                                          apply.ctypes.data_as(C.c_void_p))
    return angles, mats, apply


def deskew_decide_libm(lines_host: np.ndarray, nlines_host: np.ndarray, h: int, w: int):
    """``lumina_deskew_decide_host``: the same decision with glibc's atan2 for the per-line angles (what a host
    without numpy gets)."""
    ln = np.ascontiguousarray(lines_host, dtype=np.int32)
    nl = np.ascontiguousarray(nlines_host, dtype=np.int32)
    n, stride = ln.shape[0], ln.shape[1]
    angles = np.zeros(n, np.float64)
    mats = np.zeros((n, 6), np.float64)
    apply = np.zeros(n, np.uint8)
    _L().lumina_deskew_decide_host(ln.ctypes.data_as(C.c_void_p), nl.ctypes.data_as(C.c_void_p), n, stride, int(h), int(w),
                                   angles.ctypes.data_as(C.c_void_p), mats.ctypes.data_as(C.c_void_p),
                                   apply.ctypes.data_as(C.c_void_p))
    return angles, mats, apply


@_on_tensor_device
def warp_affine_cubic(pages: torch.Tensor, mats: np.ndarray, apply: Optional[np.ndarray] = None) -> torch.Tensor:
    """cv2.warpAffine(INTER_CUBIC, BORDER_REPLICATE) with per-page forward 2x3 matrices (host)."""
    sq = pages.dim() == 3
    x, n, h, w, c = _pages(pages)
    m = np.ascontiguousarray(mats, dtype=np.float64).reshape(n, 6)
    ap = np.ones(n, np.uint8) if apply is None else np.ascontiguousarray(apply, dtype=np.uint8)
    out = torch.empty_like(x)
    _chk(_L().lumina_warp_affine_cubic_u8(_ptr(x), _ptr(out), n, h, w, c, m.ctypes.data_as(C.c_void_p),
                                          ap.ctypes.data_as(C.c_void_p), _stream()))
    return _like(out, sq)


@_on_tensor_device
def deskew(pages: torch.Tensor, max_lines: int = 4096):
    """image_preprocessing.py:372-460 for a batch: returns (pages, angles[N] float64 numpy).

    One host synchronisation (the line lists come back for the median / gating),
    exactly where the reference computes ``np.median(angles)``."""
    x, n, h, w, c = _pages(pages)
    edges = canny(x, 50, 150)
    lines, nlines = hough_lines_p(edges, max_lines=max_lines)
    nl = nlines.cpu().numpy()
    if int(nl.max(initial=0)) > max_lines:
        return deskew(pages, max_lines=int(nl.max()))
    keep = int(nl.max(initial=0))
    lh = lines[:, :max(keep, 1)].cpu().numpy()
    angles, mats, apply = deskew_decide(lh, nl, h, w)
    if not apply.any():
        return pages, angles
    out = warp_affine_cubic(x, mats, apply)
    return (out.squeeze(-1) if pages.dim() == 3 else out), angles


@_on_tensor_device
def estimate_skew_fast(edges: torch.Tensor) -> torch.Tensor:
    """Projection-profile estimate of the reference's deskew angle from Canny edge maps [N,H,W] (degrees, f64, on the
    device).  NOT the reference's algorithm (Hough segments + median): a tolerance-certified alternative, flag-gated."""
    if not edges.is_cuda or edges.dtype != torch.uint8 or edges.dim() != 3:
        raise TypeError("estimate_skew_fast expects a CUDA uint8 [N,H,W] edge map")
    e = edges.contiguous()
    n, h, w = e.shape
    angles = torch.empty(n, dtype=torch.float64, device=e.device)
    wsb = int(_L().lumina_skew_workspace_bytes_for(n, h, w))
    ws = _ws(wsb, e.device)
    _chk(_L().lumina_skew_estimate_fast(_ptr(e), n, h, w, _ptr(angles), _ptr(ws), wsb, _stream()))
    return angles


def deskew_fast(pages: torch.Tensor):
    """The reference's deskew (image_preprocessing.py:372-460) with the angle taken from ``estimate_skew_fast``
    instead of HoughLinesP + median; gates (:433-439), rotation matrix and bicubic warp are the reference's.
    Flag-gated: rasters differ from the reference's wherever the two angles differ."""
    x, n, h, w, c = _pages(pages)
    est = estimate_skew_fast(canny(x, 50, 150)).cpu().numpy()
    angles = np.zeros(n, np.float64)
    mats = np.zeros((n, 6), np.float64)
    apply = np.zeros(n, np.uint8)
    for i in range(n):
        a = float(est[i])
        if abs(a) < 0.5:          # :433-435 (angle reported, image untouched)
            angles[i] = a
        elif abs(a) > 45:         # :437-439
            angles[i] = 0.0
        else:
            angles[i] = a
            mats[i] = rotation_matrix(w // 2, h // 2, a, 1.0).reshape(-1)
            apply[i] = 1
    if not apply.any():
        return pages, angles
    out = warp_affine_cubic(x, mats, apply)
    return (out.squeeze(-1) if pages.dim() == 3 else out), angles


# --------------------------------------------------------------------------- a15 / a17
DET_MEAN = (0.485, 0.456, 0.406)
DET_STD = (0.229, 0.224, 0.225)


DET_LIMIT_TYPES = {"max": 0, "min": 1, "resize_long": 2}


def det_target_size(h: int, w: int, limit_side_len: int = 960, limit_type: str = "max") -> Tuple[int, int]:
    """upstream DetResizeForTest.resize_image_type0: (resize_h, resize_w), multiples of 32."""
    if limit_type not in DET_LIMIT_TYPES:
        raise ValueError(f"limit_type must be one of {sorted(DET_LIMIT_TYPES)}, got {limit_type!r}")
    oh, ow = C.c_int(), C.c_int()
    _chk(_L().lumina_det_target_size_ex(int(h), int(w), int(limit_side_len), DET_LIMIT_TYPES[limit_type], C.byref(oh), C.byref(ow)))
    return oh.value, ow.value


@_on_tensor_device
def det_resize_normalize(pages: torch.Tensor, limit_side_len: int = 960, mean=DET_MEAN, std=DET_STD,
                         scale: float = 1.0 / 255.0, limit_type: str = "max"):
    """PaddleOCR DetResizeForTest(limit_side_len, limit_type) + NormalizeImage + ToCHWImage -> ([N,3,oh,ow] f32,
    shape_list).  ``limit_type``: "max" (inference default), "min" or "resize_long" (may enlarge: the same
    cv2.resize(INTER_LINEAR) arithmetic, which does not depend on the direction)."""
    x, n, h, w, c = _pages(pages)
    if c != 3:
        raise ValueError("det preprocess expects RGB/BGR pages")
    oh, ow = det_target_size(h, w, limit_side_len, limit_type)
    out = torch.empty((n, 3, oh, ow), dtype=torch.float32, device=x.device)
    m = np.asarray(mean, np.float32)
    s = np.asarray(std, np.float32)
    _chk(_L().lumina_det_resize_normalize(_ptr(x), _ptr(out), n, h, w, oh, ow, m.ctypes.data_as(C.c_void_p),
                                          s.ctypes.data_as(C.c_void_p), float(np.float32(scale)), _stream()))
    shape_list = np.tile(np.array([h, w, oh / h, ow / w], np.float64), (n, 1))
    return out, shape_list


@_on_tensor_device
def ctc_greedy(probs: torch.Tensor):
    """CTC greedy decode of [N,T,C] float32 posteriors -> (idx, pos, length, conf) device tensors."""
    if not probs.is_cuda or probs.dtype != torch.float32 or probs.dim() != 3:
        raise TypeError("ctc_greedy expects a CUDA float32 [N,T,C] tensor")
    p = probs.contiguous()
    n, t, c = p.shape
    idx = torch.empty((n, t), dtype=torch.int32, device=p.device)
    pos = torch.empty((n, t), dtype=torch.int32, device=p.device)
    ln = torch.empty(n, dtype=torch.int32, device=p.device)
    conf = torch.empty(n, dtype=torch.float32, device=p.device)
    wsb = int(_L().lumina_ctc_workspace_bytes(n, t))
    ws = _ws(wsb, p.device)
    _chk(_L().lumina_ctc_greedy(_ptr(p), n, t, c, _ptr(idx), _ptr(pos), _ptr(ln), _ptr(conf), _ptr(ws), wsb, _stream()))
    return idx, pos, ln, conf


def synth_pages(n: int, h: int = 3508, w: int = 2480, seed0: int = 0, device="cuda",
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Synthetic A4 text pages generated in HBM (identical bytes to oracle.synth_page)."""
    if out is None:
        out = torch.empty((n, h, w, 3), dtype=torch.uint8, device=device)
    with torch.cuda.device(out.device):
        _chk(_L().lumina_synth_pages_u8(_ptr(out), n, h, w, C.c_uint64(seed0), _stream()))
    return out


def synth_prob_maps(n: int, h: int = 960, w: int = 960, seed0: int = 0, device="cuda", out: Optional[torch.Tensor] = None):
    """Synthetic DB probability maps [n,h,w] f32 generated in HBM (identical floats to oracle.synth_prob_map_grid)."""
    if out is None:
        out = torch.empty((n, h, w), dtype=torch.float32, device=device)
    with torch.cuda.device(out.device):
        _chk(_L().lumina_synth_prob_maps_f32(_ptr(out), n, h, w, C.c_uint64(seed0), _stream()))
    return out


def synth_ctc(n: int, t: int = 40, c: int = 6625, crop0: int = 0, seed: int = 1, device="cuda",
              out: Optional[torch.Tensor] = None):
    """Synthetic CTC posteriors [n,t,c] f32 generated in HBM (identical floats to oracle.synth_ctc)."""
    if out is None:
        out = torch.empty((n, t, c), dtype=torch.float32, device=device)
    with torch.cuda.device(out.device):
        _chk(_L().lumina_synth_ctc_f32(_ptr(out), n, t, c, C.c_uint64(crop0), C.c_uint32(seed), _stream()))
    return out


# --------------------------------------------------------------------------- a16
@_on_tensor_device
def db_mask_ccl(pred: torch.Tensor, thresh: float = 0.3):
    """Stage outputs of the DB labelling: (mask {0,1} [N,H,W] u8, labels [N,H,W] int32 with
    label = min raster index of the 8-connected component + 1, 0 = background)."""
    if not pred.is_cuda or pred.dtype != torch.float32 or pred.dim() != 3:
        raise TypeError("db_mask_ccl expects a CUDA float32 [N,H,W] tensor")
    p = pred.contiguous()
    n, h, w = p.shape
    mask = torch.empty((n, h, w), dtype=torch.uint8, device=p.device)
    labels = torch.empty((n, h, w), dtype=torch.int32, device=p.device)
    wsb = ((n * h * w + 255) // 256) * 256 + 4 * n * h * w
    ws = _ws(wsb, p.device)
    _chk(_L().lumina_db_mask_ccl(_ptr(p), n, h, w, float(np.float32(thresh)), _ptr(mask), _ptr(labels), _ptr(ws), wsb,
                                 _stream()))
    return mask, labels


@_on_tensor_device
def db_postprocess(pred: torch.Tensor, src_hw, thresh: float = 0.3, box_thresh: float = 0.7,
                   unclip_ratio: float = 2.0, max_candidates: int = 1000, min_size: int = 3, use_dilation: bool = False,
                   score_mode: str = "fast"):
    """DBPostProcess core on [N,H,W] float32 maps -> (boxes [N,max_candidates,4,2] int32,
    scores [N,max_candidates] f32, counts [N] int32), all on the device.  ``use_dilation``: upstream's 2x2 mask
    dilation before the contours are taken.  ``score_mode``: "fast" (mean over the first quad) or "slow" (mean over
    the filled contour)."""
    if score_mode not in ("fast", "slow"):
        raise ValueError("score_mode must be 'fast' or 'slow'")
    if not pred.is_cuda or pred.dtype != torch.float32 or pred.dim() != 3:
        raise TypeError("db_postprocess expects a CUDA float32 [N,H,W] tensor")
    p = pred.contiguous()
    n, h, w = p.shape
    hw = np.ascontiguousarray(np.asarray(src_hw, dtype=np.int32).reshape(n, 2))
    boxes = torch.empty((n, max_candidates, 4, 2), dtype=torch.int32, device=p.device)
    scores = torch.empty((n, max_candidates), dtype=torch.float32, device=p.device)
    counts = torch.empty(n, dtype=torch.int32, device=p.device)
    wsb = int(_L().lumina_db_workspace_bytes(n, h, w, int(max_candidates)))
    ws = _ws(wsb, p.device)
    _chk(_L().lumina_db_postprocess_ex(_ptr(p), n, h, w, float(np.float32(thresh)), float(box_thresh), float(unclip_ratio),
                                       int(max_candidates), int(min_size), (1 if use_dilation else 0) | (2 if score_mode == "slow" else 0),
                                       hw.ctypes.data_as(C.c_void_p), _ptr(boxes), _ptr(scores), _ptr(counts), _ptr(ws), wsb,
                                       _stream()))
    return boxes, scores, counts


# --------------------------------------------------------------------------- next row 8f.2
@_on_tensor_device
def reading_order(boxes: torch.Tensor, conf: torch.Tensor, offsets: torch.Tensor, y_tolerance_ratio: float = 0.5,
                  max_boxes_per_page: Optional[int] = None, one_line: bool = False):
    """ocr_postprocessor.py:101-182 for a batch of pages.

    boxes  CUDA float64 [total,4,2], conf CUDA float64 [total], offsets int32 [pages+1] (CUDA or host).
    one_line=True: every page is one pre-grouped line (input order breaks x ties).  A negative / NaN ratio keeps
    the reference's meaning: no two blocks ever share a line.
    Returns device tensors (order[total] int32, line_of[total] int32, nlines[pages] int32,
    line_conf[total] f64, line_y[total] f64), all indexed from offsets[p]."""
    if not boxes.is_cuda or boxes.dtype != torch.float64 or boxes.dim() != 3 or tuple(boxes.shape[1:]) != (4, 2):
        raise TypeError("reading_order expects CUDA float64 boxes [total,4,2]")
    if conf.dtype != torch.float64 or not conf.is_cuda or conf.numel() != boxes.shape[0]:
        raise TypeError("reading_order expects CUDA float64 confidences [total]")
    off_host = offsets.detach().to("cpu", torch.int32)
    pages = off_host.numel() - 1
    if pages < 1 or int(off_host[0]) != 0 or int(off_host[-1]) != boxes.shape[0]:
        raise ValueError("offsets must run from 0 to the number of boxes")
    counts = (off_host[1:] - off_host[:-1])
    if int(counts.min()) < 0:
        raise ValueError("offsets must be non-decreasing")
    mb = int(counts.max()) if max_boxes_per_page is None else int(max_boxes_per_page)
    dev = boxes.device
    off_dev = off_host.to(dev)
    b = boxes.contiguous()
    c = conf.contiguous()
    total = max(int(b.shape[0]), 1)
    order = torch.empty(total, dtype=torch.int32, device=dev)
    line_of = torch.empty(total, dtype=torch.int32, device=dev)
    nlines = torch.empty(pages, dtype=torch.int32, device=dev)
    line_conf = torch.empty(total, dtype=torch.float64, device=dev)
    line_y = torch.empty(total, dtype=torch.float64, device=dev)
    if b.shape[0] == 0:
        nlines.zero_()
        return order[:0], line_of[:0], nlines, line_conf[:0], line_y[:0]
    with torch.cuda.device(dev):
        _chk(_L().lumina_reading_order(_ptr(b), _ptr(c), _ptr(off_dev), pages, mb, float(y_tolerance_ratio), int(bool(one_line)), _ptr(order),
                                       _ptr(line_of), _ptr(nlines), _ptr(line_conf), _ptr(line_y), _stream()))
    n = int(b.shape[0])
    return order[:n], line_of[:n], nlines, line_conf[:n], line_y[:n]


# --------------------------------------------------------------------------- next row 8f.1
class JpegEncoder:
    """Batched baseline JPEG encoder (image_preprocessing.py:496-557, :331-347): the byte stream Pillow writes
    for ``image.save(format='JPEG', quality=q, optimize=...)``.  Keeps its workspace (DCT coefficients of the
    last batch included) so that the quality loop of ``compress_for_azure`` re-quantises instead of
    re-transforming."""

    def __init__(self):
        self._ws = None
        self._key = None
        self._out = None      # pinned host staging [n, stride]: the D2H copies run at PCIe speed

    def encode(self, pages: torch.Tensor, quality: int = 95, optimize: bool = False, max_bytes: Optional[int] = None,
               reuse_dct: bool = False):
        """pages CUDA uint8 [N,H,W,3] -> (list of bytes-or-None, sizes int64[N]).  A page whose file is larger than
        ``max_bytes`` comes back as None (its size is still reported)."""
        x, n, h, w, c = _pages(pages)
        if c != 3:
            raise ValueError("the JPEG path encodes RGB pages (the reference converts L / RGBA / P to RGB first)")
        x = x.contiguous()
        key = (n, h, w, x.device)
        wsb = int(_L().lumina_jpeg_workspace_bytes(n, h, w))
        if self._key != key or self._ws is None:
            self._ws = _ws(wsb, x.device)
            self._key = key
            reuse_dct = False
        # without a cap: room for 1.5 bytes per pixel (quality-95 text pages need ~0.5); a page that does not fit
        # triggers one re-run with the exact size (the DCT is reused)
        stride = max(int(max_bytes), 1) if max_bytes is not None else 4096 + (3 * h * w) // 2
        sizes = np.zeros(n, np.int64)
        while True:
            if self._out is None or self._out.shape[0] < n or self._out.shape[1] < stride:
                self._out = torch.empty((n, stride), dtype=torch.uint8, pin_memory=True)
            out = self._out
            flags = (1 if optimize else 0) | (2 if reuse_dct else 0)
            with torch.cuda.device(x.device):
                _chk(_L().lumina_jpeg_encode_rgb(_ptr(x), n, h, w, int(quality), flags, C.c_void_p(out.data_ptr()),
                                                 int(out.shape[1]), sizes.ctypes.data_as(C.c_void_p), _ptr(self._ws), wsb, _stream()))
            if max_bytes is not None or int(sizes.max()) <= out.shape[1]:
                break
            stride, reuse_dct = int(sizes.max()), True
        cap = stride if max_bytes is not None else out.shape[1]
        view = out.numpy()
        files = [view[i, : int(sizes[i])].tobytes() if sizes[i] <= cap else None for i in range(n)]
        return files, sizes


def jpeg_encode(pages: torch.Tensor, quality: int = 95, optimize: bool = False):
    """One-shot form of :class:`JpegEncoder`: list of JPEG files (bytes), one per page."""
    files, sizes = JpegEncoder().encode(pages, quality, optimize)
    if any(f is None for f in files):
        raise _abi.LuminaError("a JPEG file exceeded the output buffer")
    return files


# --------------------------------------------------------------------------- next row 8f.3
class JpegInfo(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("channels", C.c_int32), ("hs", C.c_int32), ("vs", C.c_int32)]


def jpeg_probe(data) -> Optional[JpegInfo]:
    """Header parse of one file (bytes / bytearray / uint8 ndarray).  None when the file is not a JPEG the device
    decodes (progressive, CMYK, ... or not a JPEG at all): such a file stays on the host codec, which is what the
    reference uses for every file (image_preprocessing.py:63,72)."""
    buf = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else data
    info = JpegInfo()
    rc = _L().lumina_jpeg_probe(buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(info))
    return info if rc == 0 else None


class JpegDecoder:
    """Batched baseline-JPEG decode in HBM (load_image / load_image_bytes, image_preprocessing.py:57-75): the files
    cross PCIe, the rasters are produced on the device and equal ``np.asarray(Image.open(f))`` byte for byte.
    Keeps its device workspace and a ring of pinned table-staging buffers, so a stream of batches allocates
    nothing after the first call."""

    RING = 4

    def __init__(self):
        self._ws = None
        self._stage = []      # [(pinned tensor, event or None)]
        self._turn = 0
        self._blob = None     # pinned staging for callers that hand over Python bytes

    def pack(self, files):
        """list of bytes -> (pinned uint8 blob, int64 offsets [n+1]).  The blob is this decoder's own staging
        buffer: valid until the next ``pack``."""
        offs = np.zeros(len(files) + 1, np.int64)
        np.cumsum([len(f) for f in files], out=offs[1:])
        total = int(offs[-1])
        if self._blob is None or self._blob.numel() < total:
            self._blob = torch.empty(max(total, 1 << 20), dtype=torch.uint8, pin_memory=True)
        view = self._blob.numpy()
        for f, o in zip(files, offs[:-1]):
            view[o:o + len(f)] = np.frombuffer(f, np.uint8)
        return self._blob[:total], offs

    def decode(self, blob: torch.Tensor, offsets: np.ndarray, device=None, info: Optional[JpegInfo] = None):
        """blob: CPU uint8 tensor holding the files back to back (pinned memory keeps the call asynchronous),
        offsets int64 [n+1].  Returns (pages CUDA uint8 [N,H,W,C], status CUDA int32 [N]); enqueued on the
        current stream of ``device``.  The blob must stay untouched until that stream has passed this call."""
        if blob.is_cuda or blob.dtype != torch.uint8:
            raise TypeError("JpegDecoder.decode needs a CPU uint8 tensor holding the files")
        offsets = np.ascontiguousarray(offsets, np.int64)
        n = offsets.size - 1
        if n <= 0:
            raise ValueError("empty batch")
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        base = blob.numpy()
        if info is None:
            info = jpeg_probe(base[int(offsets[0]):int(offsets[1])])
            if info is None:
                raise _abi.LuminaError("page 0 is not a JPEG the device decoder covers")
        h, w, c = int(info.height), int(info.width), int(info.channels)
        total = int(offsets[-1] - offsets[0])
        with torch.cuda.device(dev):
            wsb = int(_L().lumina_jpeg_decode_workspace_bytes(n, h, w, c, info.hs, info.vs, total))
            if self._ws is None or self._ws.numel() < wsb or self._ws.device != dev:
                self._ws = None
                self._ws = torch.empty(wsb + wsb // 8, dtype=torch.uint8, device=dev)
            sb = int(_L().lumina_jpeg_decode_stage_bytes(n))
            if len(self._stage) != self.RING or self._stage[0][0].numel() < sb:
                self._stage = [[torch.empty(sb, dtype=torch.uint8, pin_memory=True), None] for _ in range(self.RING)]
            slot = self._stage[self._turn % self.RING]
            self._turn += 1
            if slot[1] is not None:
                slot[1].synchronize()   # the upload that last read this staging slot has finished
            pages = torch.empty((n, h, w, c), dtype=torch.uint8, device=dev)
            status = torch.empty(n, dtype=torch.int32, device=dev)
            _chk(_L().lumina_jpeg_decode_batch(C.c_void_p(blob.data_ptr()), offsets.ctypes.data_as(C.c_void_p), n, h, w, c,
                                               _ptr(pages), _ptr(status), C.c_void_p(slot[0].data_ptr()), _ptr(self._ws),
                                               int(self._ws.numel()), _stream()))
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            slot[1] = ev
        return pages, status


def jpeg_decode(files, device=None):
    """One-shot form: list of JPEG files (bytes) of one geometry -> CUDA uint8 [N,H,W,C] (synchronises)."""
    d = JpegDecoder()
    blob, offs = d.pack(files)
    pages, status = d.decode(blob, offs, device)
    bad = torch.nonzero(status).flatten().tolist()
    if bad:
        raise _abi.LuminaError(f"corrupt entropy-coded data in pages {bad}")
    return pages
