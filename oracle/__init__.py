"""CPU ORACLE -- test infrastructure only (never imported by the product path).

numpy-in / numpy-out ctypes wrappers over ``oracle/lumina_oracle.c`` (the plain-C
restatement of the arithmetic behind the reference's
``backend/utils/image_preprocessing.py`` call sites), plus ``oracle.db_post``
(upstream PaddleOCR DBPostProcess restated with cv2 + a restated Clipper offset).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblumina_oracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("lumina_oracle.c", "jpeg_decode.c")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_median_angle.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _hwc(img):
    img = _u8(img)
    if img.ndim == 2:
        return img, img.shape[0], img.shape[1], 1
    return img, img.shape[0], img.shape[1], img.shape[2]


def resize_nearest(plane, out_w: int, out_h: int):
    """Pillow's Image.resize on mode "P" / "1" objects (image_preprocessing.py:110 reaches it for those modes):
    the filter is forced to NEAREST; Geometry.c ImagingScaleAffine maps output x to (int)(a/2 + x*a) with the
    products accumulated by repeated addition in double.  Pure-Python table, numpy gather."""
    plane = _u8(plane)

    def tab(n_in, n_out):
        a = n_in / n_out
        xo = 0.0 + a * 0.5
        t = np.empty(n_out, np.int64)
        for x in range(n_out):
            t[x] = -1 if xo < 0.0 else int(xo)
            xo += a
        return t

    xi, yi = tab(plane.shape[1], out_w), tab(plane.shape[0], out_h)
    return plane[yi][:, xi]


# --- A1 ---------------------------------------------------------------------
def target_size(width: int, height: int, max_dim: int):
    """image_preprocessing.py:97-105 (int() truncation)."""
    if max(width, height) <= max_dim:
        return width, height
    if width > height:
        return max_dim, int(height * (max_dim / width))
    return int(width * (max_dim / height)), max_dim


def lanczos_coeffs(in_size: int, out_size: int):
    k = lib().orc_lanczos_ksize(in_size, out_size)
    b = np.zeros((out_size, 2), np.int32)
    c = np.zeros((out_size, k), np.int32)
    lib().orc_lanczos_coeffs(in_size, out_size, _p(b), _p(c))
    return b, c


def resize_lanczos(img, out_w: int, out_h: int):
    img, h, w, c = _hwc(img)
    out = np.empty((out_h, out_w) + ((c,) if img.ndim == 3 else ()), np.uint8)
    lib().orc_resize_lanczos_u8(_p(img), h, w, c, _p(out), out_h, out_w)
    return out


# --- A2 ---------------------------------------------------------------------
def gray_pil(rgb):
    rgb = _u8(rgb)
    out = np.empty(rgb.shape[:2], np.uint8)
    lib().orc_gray_pil(_p(rgb), C.c_size_t(out.size), _p(out))
    return out


def gray_cv(rgb):
    rgb = _u8(rgb)
    out = np.empty(rgb.shape[:2], np.uint8)
    lib().orc_gray_cv(_p(rgb), C.c_size_t(out.size), _p(out))
    return out


# --- A3/A4/A5 ---------------------------------------------------------------
def contrast_mean(img):
    img, h, w, c = _hwc(img)
    return int(lib().orc_contrast_mean(_p(img), C.c_size_t(h * w), c))


def contrast(img, factor: float):
    img, h, w, c = _hwc(img)
    out = np.empty_like(img)
    lib().orc_contrast(_p(img), h, w, c, C.c_float(factor), _p(out))
    return out


def smooth3(img):
    img, h, w, c = _hwc(img)
    out = np.empty_like(img)
    lib().orc_smooth3(_p(img), h, w, c, _p(out))
    return out


def sharpness(img, factor: float):
    img, h, w, c = _hwc(img)
    out = np.empty_like(img)
    lib().orc_sharpness(_p(img), h, w, c, C.c_float(factor), _p(out))
    return out


def median3(img):
    img, h, w, c = _hwc(img)
    out = np.empty_like(img)
    lib().orc_median3(_p(img), h, w, c, _p(out))
    return out


def threshold(gray, thr: int = 128):
    gray = _u8(gray)
    out = np.empty_like(gray)
    lib().orc_threshold(_p(gray), C.c_size_t(gray.size), thr, _p(out))
    return out


def exif_transpose(img, orientation: int):
    img, h, w, c = _hwc(img)
    oh, ow = (w, h) if 5 <= orientation <= 8 else (h, w)
    out = np.empty((oh, ow) + ((c,) if img.ndim == 3 else ()), np.uint8)
    lib().orc_exif_transpose(_p(img), h, w, c, orientation, _p(out))
    return out


# --- A6 ---------------------------------------------------------------------
CV_DISPATCH = {"plain": 0, "avx2": 1}


def adaptive_gauss11(gray, cval: int = 2, cv_dispatch: str = "plain"):
    """cv2.adaptiveThreshold(GAUSSIAN_C, BINARY, 11, cval).  ``cv_dispatch``: "plain" = OpenCV's plain float path
    (cv2.setUseOptimized(False): the 217 goldens), "avx2" = OpenCV's default dispatch on x86 hosts with AVX2 + FMA3
    (what the reference runs unless told otherwise; tests/golden/adaptive_dispatch_golden.json)."""
    gray = _u8(gray)
    out = np.empty_like(gray)
    lib().orc_adaptive_gauss11_x(_p(gray), gray.shape[0], gray.shape[1], cval, CV_DISPATCH[cv_dispatch], _p(out))
    return out


# --- A7/A8/A9 ---------------------------------------------------------------
def canny(gray, low: int = 50, high: int = 150):
    gray = _u8(gray)
    out = np.empty_like(gray)
    lib().orc_canny(_p(gray), gray.shape[0], gray.shape[1], low, high, _p(out))
    return out


def ppht(edges, rho=1.0, theta=np.pi / 180, threshold=100, min_len=100, max_gap=10, max_lines=1 << 16):
    edges = _u8(edges)
    lines = np.zeros((max_lines, 4), np.int32)
    n = lib().orc_ppht(_p(edges), edges.shape[0], edges.shape[1], C.c_double(rho), C.c_double(theta),
                       threshold, min_len, max_gap, _p(lines), max_lines)
    return lines[: min(n, max_lines)].copy()


def median_angle(lines) -> float:
    """image_preprocessing.py:417-428, statement by statement, WITH NUMPY: np.arctan2 is not glibc's atan2 on AVX-512
    numpy builds (bundled SIMD math; ~0.3 % of segments differ in the last place), so the C restatement
    (orc_median_angle, glibc) is only the reference's value where numpy dispatches to libm.  Pinned by the deskew
    goldens (tests/test_oracle_golden.py)."""
    lines = np.ascontiguousarray(lines, np.int32).reshape(-1, 4)
    if len(lines) == 0:
        return 0.0
    angles = []
    for x1, y1, x2, y2 in lines:
        angle = np.degrees(np.arctan2(y2 - y1, x2 - x1))
        if angle < -45:
            angle = angle + 90
        elif angle > 45:
            angle = angle - 90
        angles.append(angle)
    return float(np.median(angles))


def median_angle_libm(lines) -> float:
    """The C restatement with glibc's atan2 (see median_angle)."""
    lines = np.ascontiguousarray(lines, np.int32)
    return float(lib().orc_median_angle(_p(lines), len(lines)))


def rotation_matrix(cx: float, cy: float, angle: float, scale: float = 1.0):
    m = np.zeros(6, np.float64)
    lib().orc_rotation_matrix(C.c_double(cx), C.c_double(cy), C.c_double(angle), C.c_double(scale), _p(m))
    return m.reshape(2, 3)


def cubic_table():
    t = np.zeros((1024, 16), np.int16)
    lib().orc_cubic_table(_p(t))
    return t


def warp_affine_cubic(img, M):
    img, h, w, c = _hwc(img)
    M = np.ascontiguousarray(M, np.float64).reshape(6)
    out = np.empty_like(img)
    lib().orc_warp_affine_cubic_u8(_p(img), h, w, c, _p(M), _p(out))
    return out


def deskew(img):
    """image_preprocessing.py:372-460 composed from the restated stages.
    Returns (image, angle, lines)."""
    img = _u8(img)
    gray = img if img.ndim == 2 else gray_cv(img)
    edges = canny(gray, 50, 150)
    lines = ppht(edges)
    if len(lines) == 0:
        return img, 0.0, lines
    angle = median_angle(lines)
    if abs(angle) < 0.5:
        return img, angle, lines
    if abs(angle) > 45:
        return img, 0.0, lines
    h, w = img.shape[:2]
    M = rotation_matrix(w // 2, h // 2, angle)
    return warp_affine_cubic(img, M), angle, lines


# --- B3 / B2 ----------------------------------------------------------------
DET_MEAN = np.array([0.485, 0.456, 0.406], np.float32)
DET_STD = np.array([0.229, 0.224, 0.225], np.float32)


def det_target_size(h: int, w: int, limit: int = 960, limit_type: str = "max"):
    if limit_type != "max":
        return det_target_size_upstream(h, w, limit, limit_type)
    rh, rw = C.c_int(), C.c_int()
    lib().orc_det_target_size(h, w, limit, C.byref(rh), C.byref(rw))
    return rh.value, rw.value


def det_target_size_upstream(h: int, w: int, limit_side_len: int = 960, limit_type: str = "max"):
    """[upstream PaddleOCR, from memory; parity unpinned] ppocr/data/imaug/operators.py DetResizeForTest.resize_image_type0,
    the size arithmetic only, statement for statement (python floats, python round)."""
    if limit_type == "max":
        if max(h, w) > limit_side_len:
            ratio = float(limit_side_len) / h if h > w else float(limit_side_len) / w
        else:
            ratio = 1.0
    elif limit_type == "min":
        if min(h, w) < limit_side_len:
            ratio = float(limit_side_len) / h if h < w else float(limit_side_len) / w
        else:
            ratio = 1.0
    elif limit_type == "resize_long":
        ratio = float(limit_side_len) / max(h, w)
    else:
        raise Exception("not support limit type, image ")
    resize_h = int(h * ratio)
    resize_w = int(w * ratio)
    resize_h = max(int(round(resize_h / 32) * 32), 32)
    resize_w = max(int(round(resize_w / 32) * 32), 32)
    return resize_h, resize_w


def resize_linear(img, out_w: int, out_h: int):
    img, h, w, c = _hwc(img)
    out = np.empty((out_h, out_w) + ((c,) if img.ndim == 3 else ()), np.uint8)
    lib().orc_resize_linear_u8(_p(img), h, w, c, _p(out), out_h, out_w)
    return out


def det_resize_normalize(img, limit: int = 960, limit_type: str = "max"):
    img = _u8(img)
    h, w = img.shape[:2]
    rh, rw = det_target_size(h, w, limit, limit_type)
    r = resize_linear(img, rw, rh)
    out = np.empty((3, rh, rw), np.float32)
    lib().orc_normalize_chw(_p(r), rh, rw, _p(DET_MEAN), _p(DET_STD), C.c_float(np.float32(1.0 / 255.0)), _p(out))
    return out, (h, w, rh / h, rw / w)


def ctc_greedy(probs):
    probs = np.ascontiguousarray(probs, np.float32)
    n, t, c = probs.shape
    idx = np.empty((n, t), np.int32)
    pos = np.empty((n, t), np.int32)
    ln = np.empty(n, np.int32)
    conf = np.empty(n, np.float32)
    lib().orc_ctc_greedy(_p(probs), n, t, c, _p(idx), _p(pos), _p(ln), _p(conf))
    return idx, pos, ln, conf


# --- synthetic workload (host build of include/lumina_synth.h) ---------------
def synth_page(h: int, w: int, seed: int):
    out = np.empty((h, w, 3), np.uint8)
    lib().orc_synth_page(_p(out), h, w, C.c_uint64(seed))
    return out


def synth_skew_deg(h: int, w: int, seed: int) -> float:
    lib().orc_synth_skew_deg.restype = C.c_double
    return float(lib().orc_synth_skew_deg(h, w, C.c_uint64(seed)))


def synth_prob_map_grid(h: int = 960, w: int = 960, seed: int = 0):
    """Host build of the device generator ops.synth_prob_maps (one map, seed = seed0 + map index)."""
    out = np.empty((h, w), np.float32)
    lib().orc_synth_prob_map(_p(out), h, w, C.c_uint64(seed))
    return out


def synth_ctc(n: int, t: int = 40, c: int = 6625, crop0: int = 0, seed: int = 1):
    """Host build of the device generator ops.synth_ctc."""
    out = np.empty((n, t, c), np.float32)
    lib().orc_synth_ctc(_p(out), n, t, c, C.c_uint64(crop0), C.c_uint32(seed))
    return out


# --- ingest: baseline JPEG decode (load_image / load_image_bytes) ------------
def jpeg_decode(data: bytes):
    """image_preprocessing.py:57-75: Image.open(...) on a baseline JPEG (libjpeg-turbo islow IDCT, fancy
    upsampling, YCbCr->RGB) -> HxWx3 (or HxW for grayscale files).  None when the file is outside the
    restated subset (progressive, CMYK, 4:4:0 ...); ValueError when it is malformed."""
    buf = np.frombuffer(data, np.uint8)
    w, h, c = C.c_int(), C.c_int(), C.c_int()
    rc = lib().orc_jpeg_decode(_p(buf), C.c_size_t(buf.size), None, C.byref(w), C.byref(h), C.byref(c))
    if rc == -4:
        return None
    if rc:
        raise ValueError("malformed JPEG")
    out = np.empty((h.value, w.value, c.value), np.uint8)
    rc = lib().orc_jpeg_decode(_p(buf), C.c_size_t(buf.size), _p(out), C.byref(w), C.byref(h), C.byref(c))
    if rc:
        raise ValueError("malformed JPEG")
    return out[:, :, 0] if c.value == 1 else out


# --- north_star extras: Otsu / Sauvola (not called by the reference) ---------
def otsu_threshold(gray) -> int:
    """getThreshVal_Otsu_8u (OpenCV modules/imgproc/src/thresh.cpp) restated in float64, same operation order.
    Pinned against cv2.threshold(..., THRESH_OTSU) in tests/test_oracle_pins.py."""
    gray = _u8(gray)
    h = np.bincount(gray.ravel(), minlength=256).astype(np.float64)
    scale = 1.0 / gray.size
    mu = 0.0
    for i in range(256):
        mu += i * h[i]
    mu *= scale
    mu1 = q1 = max_sigma = 0.0
    max_val = 0
    eps = float(np.finfo(np.float32).eps)
    for i in range(256):
        p_i = h[i] * scale
        mu1 *= q1
        q1 += p_i
        q2 = 1.0 - q1
        if min(q1, q2) < eps or max(q1, q2) > 1.0 - eps:
            continue
        mu1 = (mu1 + i * p_i) / q1
        mu2 = (mu - q1 * mu1) / q2
        sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2)
        if sigma > max_sigma:
            max_sigma, max_val = sigma, i
    return int(max_val)


def sauvola(gray, window: int = 25, k: float = 0.2, r: float = 128.0):
    """Sauvola threshold with the window clipped to the page: exact integer sums (int64 integral images), then
    T = m * (1 + k * (s / r - 1)) in float64 with separate operations.  PARITY UNPINNED by the reference (it has no
    such operator) and by any library in this image (skimage is absent): this restatement defines the operator."""
    g = _u8(gray).astype(np.int64)
    hh, ww = g.shape
    rad = window // 2
    i1 = np.zeros((hh + 1, ww + 1), np.int64)
    i2 = np.zeros((hh + 1, ww + 1), np.int64)
    i1[1:, 1:] = g.cumsum(0).cumsum(1)
    i2[1:, 1:] = (g * g).cumsum(0).cumsum(1)
    ys, xs = np.arange(hh), np.arange(ww)
    y0, y1 = np.maximum(ys - rad, 0), np.minimum(ys + rad, hh - 1) + 1
    x0, x1 = np.maximum(xs - rad, 0), np.minimum(xs + rad, ww - 1) + 1

    def box(ii):
        return ii[y1[:, None], x1[None, :]] - ii[y0[:, None], x1[None, :]] - ii[y1[:, None], x0[None, :]] + ii[y0[:, None], x0[None, :]]

    cnt = ((y1 - y0)[:, None] * (x1 - x0)[None, :]).astype(np.float64)
    m = box(i1).astype(np.float64) / cnt
    var = box(i2).astype(np.float64) / cnt - m * m
    var = np.where(var < 0.0, 0.0, var)
    s = np.sqrt(var)
    t = m * (1.0 + k * (s / r - 1.0))
    return np.where(g.astype(np.float64) > t, 255, 0).astype(np.uint8)
