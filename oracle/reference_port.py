"""CPU ORACLE, library level -- test / baseline infrastructure only.

Port of the reference's call sequence (backend/utils/image_preprocessing.py) onto
the same third-party native libraries the reference uses (Pillow, OpenCV, NumPy;
un-pinned in requirements.txt:20-23 -- the versions are whatever the image has:
Pillow 12.2.0, opencv-python-headless 4.13.0).  It is what `bench.py --impl
reference` and `cpu_baseline` time (kind "port": /root/reference itself does not
travel to the GPU box), and a second checker next to oracle/lumina_oracle.c.

Each function cites the reference lines it restates.  Nothing under
ocr-system_b200/ imports this module.
"""
from __future__ import annotations

import numpy as np

try:  # the reference treats OpenCV as optional (image_preprocessing.py:21-26)
    import cv2

    CV2_AVAILABLE = True
except ImportError:  # pragma: no cover
    cv2 = None
    CV2_AVAILABLE = False
from PIL import Image, ImageEnhance, ImageFilter, ImageOps


_REAL = None


def real_preprocessor():
    """The reference's own ImagePreprocessor from oracle/_ref/ (placed there by oracle/make_ref.py when the reference
    tree is present), or None.  When it exists the timed page chain below calls ITS methods (kind "reference")."""
    global _REAL
    if _REAL is None:
        import importlib.util
        import logging
        import os
        import sys

        here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
        path = os.path.join(here, "image_preprocessing.py")
        if not os.path.exists(path):
            _REAL = False
        else:
            saved = sys.modules.get("config")
            spec_c = importlib.util.spec_from_file_location("config", os.path.join(here, "config.py"))
            cfg = importlib.util.module_from_spec(spec_c)
            spec_c.loader.exec_module(cfg)
            sys.modules["config"] = cfg
            try:
                spec = importlib.util.spec_from_file_location("_lumina_reference_image_preprocessing", path)
                mod = importlib.util.module_from_spec(spec)
                sys.dont_write_bytecode = True
                spec.loader.exec_module(mod)
                logging.getLogger(mod.__name__).setLevel(logging.WARNING)   # the module logs every call at INFO
                mod.logger.setLevel(logging.WARNING)
                _REAL = mod
            finally:
                if saved is not None:
                    sys.modules["config"] = saved
                else:
                    sys.modules.pop("config", None)
    return _REAL or None


def kind() -> str:
    return "reference" if real_preprocessor() is not None else "port"


def resize_if_needed(image: Image.Image, max_dim: int) -> Image.Image:
    """:81-110"""
    width, height = image.size
    if max(width, height) <= max_dim:
        return image
    if width > height:
        new_size = (max_dim, int(height * (max_dim / width)))
    else:
        new_size = (int(width * (max_dim / height)), max_dim)
    return image.resize(new_size, Image.Resampling.LANCZOS)


def enhance_contrast(image, factor=1.3):
    """:132-144"""
    return ImageEnhance.Contrast(image).enhance(factor)


def enhance_sharpness(image, factor=1.2):
    """:146-158"""
    return ImageEnhance.Sharpness(image).enhance(factor)


def denoise(image):
    """:160-165"""
    return image.filter(ImageFilter.MedianFilter(size=3))


def convert_to_grayscale(image):
    """:167-169"""
    return image.convert("L")


def auto_orient(image):
    """:171-173"""
    return ImageOps.exif_transpose(image)


def binarize(image, threshold=128):
    """:175-185"""
    return image.convert("L").point(lambda x: 255 if x > threshold else 0, "1")


def deskew(image: Image.Image):
    """:372-460 -> (image, angle)"""
    if image.mode == "L":
        cv_image = np.array(image)
        gray = cv_image
    else:
        cv_image = cv2.cvtColor(np.array(image.convert("RGB")), cv2.COLOR_RGB2BGR)
        gray = cv2.cvtColor(cv_image, cv2.COLOR_BGR2GRAY)
    edges = cv2.Canny(gray, 50, 150, apertureSize=3)
    lines = cv2.HoughLinesP(edges, 1, np.pi / 180, threshold=100, minLineLength=100, maxLineGap=10)
    if lines is None:
        return image, 0.0
    angles = []
    for line in lines:
        x1, y1, x2, y2 = line[0]
        angle = np.degrees(np.arctan2(y2 - y1, x2 - x1))
        if angle < -45:
            angle = angle + 90
        elif angle > 45:
            angle = angle - 90
        angles.append(angle)
    angle = float(np.median(angles))
    if abs(angle) < 0.5:
        return image, angle
    if abs(angle) > 45:
        return image, 0.0
    h, w = cv_image.shape[:2]
    M = cv2.getRotationMatrix2D((w // 2, h // 2), angle, 1.0)
    out = cv2.warpAffine(cv_image, M, (w, h), flags=cv2.INTER_CUBIC, borderMode=cv2.BORDER_REPLICATE)
    if image.mode == "L":
        return Image.fromarray(out), angle
    return Image.fromarray(cv2.cvtColor(out, cv2.COLOR_BGR2RGB)), angle


def adaptive_binarize(image):
    """:462-494"""
    gray = np.array(image.convert("L")) if image.mode != "L" else np.array(image)
    return Image.fromarray(cv2.adaptiveThreshold(gray, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, 11, 2))


def det_resize_normalize(img: np.ndarray, limit: int = 960):
    """[upstream PaddleOCR] DetResizeForTest('max') + NormalizeImage + ToCHWImage."""
    h, w = img.shape[:2]
    ratio = float(limit) / max(h, w) if max(h, w) > limit else 1.0
    rh, rw = int(h * ratio), int(w * ratio)
    rh, rw = max(int(round(rh / 32) * 32), 32), max(int(round(rw / 32) * 32), 32)
    r = cv2.resize(img, (rw, rh))
    mean = np.array([0.485, 0.456, 0.406], np.float32).reshape(1, 1, 3)
    std = np.array([0.229, 0.224, 0.225], np.float32).reshape(1, 1, 3)
    x = (r.astype("float32") * np.float32(1.0 / 255.0) - mean) / std
    return x.transpose(2, 0, 1), (h, w, rh / h, rw / w)


def load_image_bytes(image_bytes: bytes) -> Image.Image:
    """image_preprocessing.py:70-75."""
    import io

    image = Image.open(io.BytesIO(image_bytes))
    if image.mode not in ('RGB', 'L'):
        image = image.convert('RGB')
    return image


def page_chain(page, max_dim: int = 960, enhance: bool = False):
    """The bench workload (BASELINE.json configs[1]) on one page, reference calls only:
    [load_image_bytes when `page` is an encoded file] -> resize -> deskew -> [contrast 1.2, sharpness 1.1] ->
    gray -> adaptive binarize -> det normalize.
    Returns (deskewed RGB u8, angle, gray u8, binary u8, normalized CHW f32)."""
    real = real_preprocessor()
    if real is not None:   # the reference's own code (oracle/_ref), same call sequence
        ip = real.ImagePreprocessor(max_dimension=max_dim)
        img = ip.load_image_bytes(page) if isinstance(page, (bytes, bytearray)) else Image.fromarray(page)
        img = ip.resize_if_needed(img)
        img, angle = ip.deskew(img)
        if enhance:
            img = ip.enhance_sharpness(ip.enhance_contrast(img, 1.2), 1.1)
        gray = ip.convert_to_grayscale(img)
        binary = ip.adaptive_binarize(img)
        rgb = np.asarray(img)
        norm, _ = det_resize_normalize(rgb, 960)   # upstream PaddleOCR op: not in the reference
        return rgb, angle, np.asarray(gray), np.asarray(binary), norm
    img = load_image_bytes(page) if isinstance(page, (bytes, bytearray)) else Image.fromarray(page)
    img = resize_if_needed(img, max_dim)
    img, angle = deskew(img)
    if enhance:
        img = enhance_sharpness(enhance_contrast(img, 1.2), 1.1)
    gray = convert_to_grayscale(img)
    binary = adaptive_binarize(img)
    rgb = np.asarray(img)
    norm, _ = det_resize_normalize(rgb, 960)
    return rgb, angle, np.asarray(gray), np.asarray(binary), norm


def compress_for_azure(image: Image.Image, target_size_mb: float = 2.0, initial_quality: int = 95, min_quality: int = 30) -> bytes:
    """:496-557 -- JPEG quality ladder (optimize=True), then Lanczos shrink by sqrt(target/current)."""
    import io

    target_bytes = int(target_size_mb * 1024 * 1024)
    if image.mode in ("RGBA", "P"):
        image = image.convert("RGB")
    elif image.mode == "L":
        image = image.convert("RGB")
    quality = initial_quality
    while quality >= min_quality:
        buffer = io.BytesIO()
        image.save(buffer, format="JPEG", quality=quality, optimize=True)
        if buffer.tell() <= target_bytes:
            return buffer.getvalue()
        quality -= 10
    buffer = io.BytesIO()
    image.save(buffer, format="JPEG", quality=min_quality)
    scale = (target_bytes / buffer.tell()) ** 0.5
    resized = image.resize((int(image.width * scale), int(image.height * scale)), Image.Resampling.LANCZOS)
    buffer = io.BytesIO()
    resized.save(buffer, format="JPEG", quality=min_quality, optimize=True)
    return buffer.getvalue()


def preprocess_for_azure(page: np.ndarray, max_dim: int = 2000, apply_deskew: bool = True, apply_binarize: bool = False,
                         target_size_mb: float = 2.0) -> bytes:
    """:559-626 -- what OCRService._process_single_image_sync runs per page before the Azure call:
    auto_orient -> resize_if_needed -> deskew -> (adaptive_binarize | contrast 1.2 -> sharpness 1.1) -> compress."""
    img = auto_orient(Image.fromarray(page))
    img = resize_if_needed(img, max_dim)
    if apply_deskew:
        img, _ = deskew(img)
    if apply_binarize:
        img = adaptive_binarize(img)
    else:
        img = enhance_sharpness(enhance_contrast(img, 1.2), 1.1)
    return compress_for_azure(img, target_size_mb)


_PAGES = None  # inherited by the forked workers: no per-task pickling of 26 MB rasters


def _worker_init():
    if CV2_AVAILABLE:
        cv2.setNumThreads(1)


def _worker_chain(args):
    idx, max_dim, enhance = args
    out = page_chain(_PAGES[idx], max_dim, enhance)
    return float(out[1])


def _worker_azure(args):
    idx, max_dim = args
    return preprocess_for_azure(_PAGES[idx], max_dim)


def run_pool_azure(pages, max_dim: int = 2000, procs: int | None = None):
    """preprocess_for_azure over `pages`, one page per task on `procs` workers -> (seconds, list of JPEG bytes)."""
    import multiprocessing as mp
    import os
    import time

    global _PAGES
    procs = procs or os.cpu_count() or 1
    _PAGES = pages
    try:
        with mp.get_context("fork").Pool(procs, initializer=_worker_init) as pool:
            pool.map(_worker_azure, [(i, max_dim) for i in range(min(len(pages), procs))])   # warm the workers
            t0 = time.perf_counter()
            out = pool.map(_worker_azure, [(i, max_dim) for i in range(len(pages))])
            dt = time.perf_counter() - t0
    finally:
        _PAGES = None
    return dt, out


def run_pool(pages, max_dim: int = 960, enhance: bool = False, procs: int | None = None):
    """Time the chain over `pages` with one page per task on `procs` worker processes
    (cv2 single-threaded per worker, rasters shared with the workers by fork).
    Returns (seconds, angles)."""
    import multiprocessing as mp
    import os
    import time

    global _PAGES
    procs = procs or os.cpu_count() or 1
    _PAGES = pages
    ctx = mp.get_context("fork")
    with ctx.Pool(procs, initializer=_worker_init) as pool:
        pool.map(_worker_chain, [(0, max_dim, enhance)] * procs)  # warm the workers
        t0 = time.perf_counter()
        angles = pool.map(_worker_chain, [(i, max_dim, enhance) for i in range(len(pages))], chunksize=1)
        dt = time.perf_counter() - t0
    _PAGES = None
    return dt, angles
