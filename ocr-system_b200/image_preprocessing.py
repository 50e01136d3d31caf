"""Drop-in ``ImagePreprocessor`` backed by the sm_100a kernels.

Mirrors the public surface of the reference's
``backend/utils/image_preprocessing.py`` (class ``ImagePreprocessor`` :35, singleton
``image_preprocessor`` :632): same method names, argument meaning, return types
(PIL in -> PIL out, ``deskew`` -> ``(image, angle)``, ``preprocess_for_azure`` ->
JPEG bytes) and error behaviour (``FileNotFoundError`` :61,277, ``ImportError``
:268; ``deskew`` degrades to ``(image, 0.0)`` instead of raising).  Every pixel
operation runs on the GPU through the C-ABI; codecs (PNG/JPEG/PDF) stay on the
host exactly as in the reference -- they are the boundary, not the path.

A single call is a batch of one (H2D + kernels + D2H); throughput comes from the
batched API in ``pipeline.PagePipeline``.
"""
from __future__ import annotations

import io
from concurrent.futures import ThreadPoolExecutor
import logging
import os
import threading
from pathlib import Path
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
from PIL import Image

from . import ops

logger = logging.getLogger(__name__)

ImageSource = Union[str, Path, Image.Image, bytes]


def _setting(name: str, default):
    """Hot-path knobs of the reference's ``config.settings`` (config.py:69,85-87).

    Inside the reference app its own ``config.settings`` object is honoured; standalone
    the same names are read from the environment."""
    try:  # drop-in inside the reference backend: `config` is importable there
        from config import settings  # type: ignore

        return getattr(settings, name, default)
    except Exception:
        raw = os.environ.get(name)
        if raw is None:
            return default
        if isinstance(default, bool):
            return raw.strip().lower() in ("1", "true", "yes", "on")
        return type(default)(raw)


class ImagePreprocessor:
    """GPU image preprocessing with the reference's call signatures."""

    def __init__(self, max_dimension: int = None, target_dpi: int = 300, device: Optional[str] = None,
                 cv_dispatch: Optional[str] = None):
        self.max_dimension = max_dimension or _setting("OCR_MAX_IMAGE_DIMENSION", 2000)
        self.target_dpi = target_dpi
        self._device = device
        # which OpenCV build adaptive_binarize reproduces: None = ops.default_cv_dispatch() ("avx2": OpenCV's default
        # dispatch, what the reference runs), "plain" = cv2.setUseOptimized(False)
        self.cv_dispatch = cv_dispatch
        # The reference singleton is stateless and is called from asyncio.to_thread workers
        # (services/ocr_service.py:674-676).  This one owns a pinned staging buffer and a JPEG workspace: both are
        # per calling thread, so concurrent callers never see each other's pixels or DCT coefficients.
        self._tls = threading.local()
        self._raise_pil_block_size()   # images opened / rasterised from now on qualify for the zero-copy ingest

    # ------------------------------------------------------------------ plumbing
    @property
    def device(self) -> torch.device:
        if not torch.cuda.is_available():
            raise RuntimeError("ocr_system_b200 needs a CUDA device: there is no CPU fallback for the pixel path")
        return torch.device(self._device or f"cuda:{torch.cuda.current_device()}")

    # ------------------------------------------------------------------ ingest
    # np.asarray(image) runs Pillow's raw encoder (7-10 ms per A4 page, holds the GIL).  Pillow also exports its own
    # storage through the Arrow C data interface without a copy -- mode "RGB" as 4 bytes per pixel (R, G, B, pad),
    # mode "L" as bytes -- when the image lives in ONE allocator block; a plain (GIL-free, threaded) copy of that
    # view into the pinned staging buffer is ~0.5 ms per page, and the pad byte is dropped on the device.  Pillow's
    # block size (16 MB by default) is raised once so that A4 300-dpi pages (35 MB as RGBX) qualify; images that
    # were allocated earlier, other modes or a missing pyarrow simply take the np.asarray path.
    _BLOCK_MB = int(os.environ.get("LUMINA_PIL_BLOCK_MB", "64"))
    # pages of one (size, mode) group that preprocess_pages_for_azure sends through the device at once
    _MAX_BATCH_PAGES = max(1, int(os.environ.get("LUMINA_MAX_BATCH_PAGES", "64")))

    @classmethod
    def _raise_pil_block_size(cls) -> None:
        try:
            if Image.core.get_block_size() < cls._BLOCK_MB << 20:
                Image.core.set_block_size(cls._BLOCK_MB << 20)
        except Exception:  # noqa: BLE001 - a tuning knob, never an error
            pass

    @staticmethod
    def _zero_copy_view(image: Image.Image) -> Optional[np.ndarray]:
        """[H,W,4] (mode RGB) or [H,W] (mode L) view of Pillow's own pixel storage, or None."""
        try:
            import pyarrow as pa

            image.load()
            arr = pa.array(image)
            w, h = image.size
            if image.mode == "RGB":
                v = arr.flatten().to_numpy(zero_copy_only=True)
                return v.reshape(h, w, 4) if v.size == h * w * 4 else None
            if image.mode == "L":
                v = arr.to_numpy(zero_copy_only=True)
                return v.reshape(h, w) if v.size == h * w else None
        except Exception:  # noqa: BLE001 - multi-block image, unsupported mode, no pyarrow: fall back
            return None
        return None

    def _upload(self, images: Sequence[Image.Image]) -> torch.Tensor:
        """Same-size, same-mode (RGB or L) PIL images -> one resident batch [N,H,W,C]."""
        w, h = images[0].size
        mode = images[0].mode
        if mode not in ("RGB", "L"):
            raise ValueError(f"unsupported image mode {mode!r}: load_image() converts to RGB/L first")
        n = len(images)
        views = [self._zero_copy_view(im) for im in images]
        fast = all(v is not None for v in views)
        c_mem = (4 if mode == "RGB" else 1) if fast else (3 if mode == "RGB" else 1)
        stage = self._host_stage(n, h, w, c_mem)
        sview = stage.numpy()

        def put(j):
            if fast:
                np.copyto(sview[j], views[j].reshape(h, w, c_mem))          # memcpy, releases the GIL
            else:
                sview[j] = np.asarray(images[j]).reshape(h, w, c_mem)

        if n == 1:
            put(0)
        else:
            with ThreadPoolExecutor(max_workers=min(8, n)) as ex:
                list(ex.map(put, range(n)))
        x = stage[:n].to(self.device, non_blocking=True)
        if fast and mode == "RGB":
            x = ops.rgbx_to_rgb(x)
        torch.cuda.current_stream(x.device).synchronize()   # the staging buffer is reused by the next call
        return x

    def _to_device(self, image: Image.Image) -> torch.Tensor:
        if image.mode not in ("RGB", "L"):
            raise ValueError(f"unsupported image mode {image.mode!r}: load_image() converts to RGB/L first")
        x = self._upload([image])
        if x.dim() == 3:  # L -> [1,H,W,1]: every batch on the device is NHWC
            x = x.unsqueeze(-1)
        return x

    # ---- PIL modes beyond RGB / L ---------------------------------------------------------------------------
    # The reference hands whatever PIL object it gets to Pillow, so its methods also work on RGBA / LA / RGBX /
    # CMYK / YCbCr / HSV objects (8-bit bands; Pillow processes them band by band) and, where the method starts
    # with an explicit ``image.convert(...)`` (grayscale, binarize, adaptive_binarize, deskew, compress), on every
    # mode.  Here the bands of such an image travel as a batch of single-channel planes [bands,H,W,1] through the
    # same kernels; mode changes themselves (palette lookup, CMYK->RGB, alpha premultiply) stay Pillow's, on the
    # host, exactly like ``load_image``'s ``convert('RGB')``.  Modes Pillow itself refuses for an operator (P, 1,
    # I, F ... in ImageEnhance / ImageFilter) raise the same ValueError.
    _BANDED = ("RGBA", "LA", "RGBX", "CMYK", "YCbCr", "HSV", "RGBa", "La")

    def _planes_to_device(self, image: Image.Image) -> torch.Tensor:
        a = np.asarray(image)
        planes = np.ascontiguousarray(np.moveaxis(a, -1, 0))[..., None]
        return torch.from_numpy(planes).to(self.device)

    @staticmethod
    def _planes_to_pil(t: torch.Tensor, mode: str) -> Image.Image:
        a = np.ascontiguousarray(np.moveaxis(t.squeeze(-1).cpu().numpy(), 0, -1))
        return Image.frombytes(mode, (a.shape[1], a.shape[0]), a.tobytes())

    def _check_banded(self, image: Image.Image, what: str) -> None:
        if image.mode not in self._BANDED:
            # Pillow: ImageEnhance / ImageFilter on P, 1, I, F, I;16 ... -> ValueError("image has wrong mode")
            raise ValueError(f"image has wrong mode ({image.mode!r} is not supported by {what}; "
                             "load_image() converts to RGB/L first)")

    def _gray_source(self, image: Image.Image) -> torch.Tensor:
        """Device input whose PIL-L conversion equals ``image.convert('L')`` (reference :169,:184,:481)."""
        if image.mode in ("RGB", "L"):
            return self._to_device(image)
        if image.mode == "YCbCr":                      # Pillow: YCbCr -> L is the Y band
            return self._to_device(image.getchannel(0))
        # every other mode: convert('L') == convert('RGB').convert('L') (same L24 weights; pinned per mode by
        # tests/test_abi_and_host.py::test_gray_source_rule_holds_for_every_pil_mode), so the mode change is Pillow's and the grayscale is the kernel's
        return self._to_device(image.convert("RGB"))

    @staticmethod
    def _to_pil(t: torch.Tensor) -> Image.Image:
        a = t[0]
        if a.dim() == 3 and a.shape[-1] == 1:
            a = a.squeeze(-1)
        return Image.fromarray(a.cpu().numpy())

    def _open(self, image: ImageSource) -> Image.Image:
        if isinstance(image, bytes):
            return self.load_image_bytes(image)
        if isinstance(image, (str, Path)):
            return self.load_image(image)
        return image

    # ------------------------------------------------------------------ loading (host codecs)
    def load_image(self, image_path: Union[str, Path]) -> Image.Image:
        path = Path(image_path)
        if not path.exists():
            raise FileNotFoundError(f"Image not found: {path}")
        image = Image.open(path)
        return image if image.mode in ("RGB", "L") else image.convert("RGB")

    def load_image_bytes(self, image_bytes: bytes) -> Image.Image:
        image = Image.open(io.BytesIO(image_bytes))
        return image if image.mode in ("RGB", "L") else image.convert("RGB")

    # ------------------------------------------------------------------ size
    def get_optimal_size(self, width: int, height: int, max_dimension: int = None) -> Tuple[int, int]:
        return ops.target_size(width, height, max_dimension or self.max_dimension)

    def resize_if_needed(self, image: Image.Image, max_dimension: int = None) -> Image.Image:
        max_dim = max_dimension or self.max_dimension
        width, height = image.size
        if max(width, height) <= max_dim:
            return image
        new_w, new_h = ops.target_size(width, height, max_dim)
        logger.info(f"Resizing image from {width}x{height} to {new_w}x{new_h}")
        if new_w < 1 or new_h < 1:    # Pillow's Image.resize raises this for every mode (reference :110)
            raise ValueError("height and width must be > 0")
        if image.mode in ("RGB", "L"):
            return self._to_pil(ops.resize_lanczos(self._to_device(image), new_w, new_h))
        if image.mode in ("P", "1"):      # Pillow's Image.resize forces NEAREST for these two modes
            idx = image if image.mode == "P" else image.convert("L")        # palette indices / 0-255 bytes
            plane = torch.from_numpy(np.frombuffer(bytearray(idx.tobytes()), np.uint8).reshape(1, height, width)).to(self.device)
            raw = ops.resize_nearest(plane, new_w, new_h)[0].cpu().numpy().tobytes()
            if image.mode == "1":
                return Image.frombytes("L", (new_w, new_h), raw).convert("1", dither=Image.Dither.NONE)
            out = Image.frombytes("P", (new_w, new_h), raw)
            out.putpalette(image.getpalette(image.palette.mode), image.palette.mode)
            if "transparency" in image.info:
                out.info["transparency"] = image.info["transparency"]
            return out
        self._check_banded(image, "resize")
        # Pillow's Image.resize: RGBA / LA are resampled premultiplied (RGBa / La) and converted back
        pre = {"RGBA": "RGBa", "LA": "La"}.get(image.mode)
        src = image.convert(pre) if pre else image
        out = self._planes_to_pil(ops.resize_lanczos(self._planes_to_device(src), new_w, new_h), src.mode)
        return out.convert(image.mode) if pre else out

    # ------------------------------------------------------------------ enhancement
    def enhance_contrast(self, image: Image.Image, factor: float = 1.3) -> Image.Image:
        if image.mode in ("RGB", "L"):
            return self._to_pil(ops.enhance_contrast(self._to_device(image), factor))
        # ImageEnhance.Contrast: mean of convert('L'); degenerate = that grey in the image's mode (alpha kept)
        self._check_banded(image, "ImageEnhance.Contrast")
        if image.mode in ("RGBa", "La"):
            raise ValueError(f"conversion from L to {image.mode} not supported")      # Pillow's own error
        mean = int(ops.contrast_mean(self._gray_source(image)).cpu()[0])
        grey = Image.new("L", (1, 1), mean).convert(image.mode).getpixel((0, 0))
        x = self._planes_to_device(image)
        keep = [i for i, b in enumerate(image.getbands()) if b != "A"]
        m = torch.tensor([grey[i] for i in keep], dtype=torch.int32, device=x.device)
        x[keep] = ops.enhance_contrast(x[keep].contiguous(), factor, mean=m)
        return self._planes_to_pil(x, image.mode)

    def enhance_sharpness(self, image: Image.Image, factor: float = 1.2) -> Image.Image:
        if image.mode in ("RGB", "L"):
            return self._to_pil(ops.enhance_sharpness(self._to_device(image), factor))
        self._check_banded(image, "ImageEnhance.Sharpness")
        x = self._planes_to_device(image)
        keep = [i for i, b in enumerate(image.getbands()) if b != "A"]   # degenerate.putalpha(image alpha)
        x[keep] = ops.enhance_sharpness(x[keep].contiguous(), factor)
        return self._planes_to_pil(x, image.mode)

    def denoise(self, image: Image.Image) -> Image.Image:
        if image.mode in ("RGB", "L"):
            return self._to_pil(ops.median3(self._to_device(image)))
        if image.mode == "P":
            raise ValueError("cannot filter palette images")                          # Pillow's own error
        if image.mode == "1":             # RankFilter runs on the 0/255 bytes of a bilevel image
            return self._to_pil(ops.median3(self._to_device(image.convert("L")))).convert("1", dither=Image.Dither.NONE)
        self._check_banded(image, "ImageFilter.MedianFilter")
        return self._planes_to_pil(ops.median3(self._planes_to_device(image)), image.mode)

    def convert_to_grayscale(self, image: Image.Image) -> Image.Image:
        if image.mode == "L":
            return image.copy()
        return self._to_pil(ops.gray_pil(self._gray_source(image)))

    def auto_orient(self, image: Image.Image) -> Image.Image:
        orientation = image.getexif().get(0x0112)
        if orientation not in (2, 3, 4, 5, 6, 7, 8):
            return image.copy()
        if image.mode in ("RGB", "L"):
            return self._to_pil(ops.exif_transpose(self._to_device(image), int(orientation)))
        if image.mode in self._BANDED:
            return self._planes_to_pil(ops.exif_transpose(self._planes_to_device(image), int(orientation)), image.mode)
        raise ValueError(f"unsupported image mode {image.mode!r}: load_image() converts to RGB/L first")

    def binarize(self, image: Image.Image, threshold: int = 128) -> Image.Image:
        mask = ops.binarize(self._gray_source(image), threshold)
        return Image.fromarray(mask[0].cpu().numpy()).convert("1", dither=Image.Dither.NONE)

    # ------------------------------------------------------------------ composed pipelines
    def optimize_for_ocr(self, image: ImageSource, apply_contrast: bool = True, apply_sharpness: bool = True,
                         apply_denoise: bool = False, grayscale: bool = False) -> Image.Image:
        """auto_orient -> resize -> [gray] -> [median] -> contrast 1.2 -> sharpness 1.1, one H2D/D2H."""
        img = self.auto_orient(self._open(image))
        if img.mode not in ("RGB", "L"):   # any other PIL object: the same steps, one drop-in method at a time
            img = self.resize_if_needed(img)
            if grayscale:
                img = self.convert_to_grayscale(img)
            if apply_denoise:
                img = self.denoise(img)
            if apply_contrast:
                img = self.enhance_contrast(img, factor=1.2)
            if apply_sharpness:
                img = self.enhance_sharpness(img, factor=1.1)
            return img
        x = self._to_device(img)
        x = ops.resize_if_needed(x, self.max_dimension)
        if grayscale and x.shape[-1] == 3:
            x = ops.gray_pil(x).unsqueeze(-1)
        if apply_denoise:
            x = ops.median3(x)
        if apply_contrast and apply_sharpness:
            x = ops.contrast_sharpness(x, 1.2, 1.1)
        elif apply_contrast:
            x = ops.enhance_contrast(x, 1.2)
        elif apply_sharpness:
            x = ops.enhance_sharpness(x, 1.1)
        return self._to_pil(x)

    # ------------------------------------------------------------------ PDF (host rasteriser, then GPU resize)
    def pdf_to_images(self, pdf_path: Union[str, Path], dpi: int = None) -> List[Image.Image]:
        try:
            from pdf2image import convert_from_path
        except ImportError:
            raise ImportError(
                "pdf2image not installed. Install with: pip install pdf2image\n"
                "Also install poppler: https://github.com/oschwartz10612/poppler-windows/releases"
            )
        path = Path(pdf_path)
        if not path.exists():
            raise FileNotFoundError(f"PDF not found: {path}")
        pages = convert_from_path(str(path), dpi=dpi or self.target_dpi, fmt="png")
        return self.resize_pages(pages)

    def resize_pages(self, pages: List[Image.Image]) -> List[Image.Image]:
        """The per-page ``resize_if_needed`` loop of pdf_to_images (:287-292), batched by page shape."""
        out: List[Optional[Image.Image]] = [None] * len(pages)
        groups = {}
        for i, pg in enumerate(pages):
            if pg.mode not in ("RGB", "L"):
                pg = pages[i] = pg.convert("RGB")
            if max(pg.size) <= self.max_dimension:
                out[i] = pg
            else:
                groups.setdefault((pg.size, pg.mode), []).append(i)
        cap = self._MAX_BATCH_PAGES      # a long document goes through in slices: host / HBM copies stay bounded
        for (size, _mode), members in groups.items():
            nw, nh = ops.target_size(size[0], size[1], self.max_dimension)
            for k0 in range(0, len(members), cap):
                idxs = members[k0:k0 + cap]
                batch = torch.from_numpy(np.stack([np.asarray(pages[i]) for i in idxs])).to(self.device)
                res = ops.resize_lanczos(batch, nw, nh).cpu().numpy()
                for k, i in enumerate(idxs):
                    out[i] = Image.fromarray(res[k])
        return out  # type: ignore[return-value]

    def get_pdf_page_count(self, pdf_path: Union[str, Path]) -> int:
        try:
            from pdf2image import pdfinfo_from_path

            return pdfinfo_from_path(str(pdf_path)).get("Pages", 1)
        except Exception:
            return len(self.pdf_to_images(pdf_path))

    # ------------------------------------------------------------------ utilities (host codecs)
    def save_image(self, image: Image.Image, output_path: Union[str, Path], quality: int = 95,
                   optimize: bool = True) -> Path:
        path = Path(output_path)
        path.parent.mkdir(parents=True, exist_ok=True)
        suffix = path.suffix.lower()
        if suffix in (".jpg", ".jpeg"):
            image.save(path, "JPEG", quality=quality, optimize=optimize)
        elif suffix == ".png":
            image.save(path, "PNG", optimize=optimize)
        else:
            image.save(path)
        return path

    def image_to_bytes(self, image: Image.Image, format: str = "PNG", quality: int = 95) -> bytes:
        buf = io.BytesIO()
        if format.upper() in ("JPG", "JPEG"):
            image.save(buf, format="JPEG", quality=quality)
        else:
            image.save(buf, format=format)
        return buf.getvalue()

    def get_image_info(self, image: Union[str, Path, Image.Image]) -> dict:
        img = self.load_image(image) if isinstance(image, (str, Path)) else image
        return {
            "width": img.width,
            "height": img.height,
            "mode": img.mode,
            "format": img.format,
            "size_optimal": self.get_optimal_size(img.width, img.height),
            "needs_resize": max(img.width, img.height) > self.max_dimension,
        }

    # ------------------------------------------------------------------ Azure preprocessing
    def deskew(self, image: Image.Image) -> Tuple[Image.Image, float]:
        """Canny -> HoughLinesP -> median angle -> bicubic rotation (reference :372-460).
        Never raises for a bad page: degrades to ``(image, 0.0)`` like the reference."""
        x = self._to_device(image.convert("RGB") if image.mode not in ("RGB", "L") else image)
        out, angles = ops.deskew(x)
        angle = float(angles[0])
        if out is x or out.data_ptr() == x.data_ptr():
            return image, angle
        return self._to_pil(out), angle

    def adaptive_binarize(self, image: Image.Image) -> Image.Image:
        return self._to_pil(ops.adaptive_binarize(self._gray_source(image), 2, self.cv_dispatch))

    def compress_pages_for_azure(self, pages: torch.Tensor, target_size_mb: float = 2.0, initial_quality: int = 95,
                                 min_quality: int = 30) -> List[bytes]:
        """``compress_for_azure`` (reference :496-557) for a resident batch [N,H,W,3] (or [N,H,W,1], replicated to
        RGB like ``image.convert('RGB')``): the quality ladder initial..min step 10 runs on the GPU encoder, which
        writes the byte stream Pillow would (``optimize=True``); the DCT is computed once per batch and
        re-quantised per quality.  Pages that still exceed the target at ``min_quality`` are Lanczos-shrunk by
        sqrt(target/current) and encoded once more, as in the reference."""
        target = int(target_size_mb * 1024 * 1024)
        if pages.dim() == 3:
            pages = pages.unsqueeze(-1)
        if pages.shape[-1] == 1:
            pages = pages.expand(-1, -1, -1, 3)
        cur = pages.contiguous()
        n = cur.shape[0]
        out: List[Optional[bytes]] = [None] * n
        todo = list(range(n))
        enc = self._jpeg_encoder()
        quality, fresh = initial_quality, True
        while quality >= min_quality and todo:
            files, _ = enc.encode(cur, quality, optimize=True, max_bytes=target, reuse_dct=not fresh)
            fresh = False
            keep = []
            for j, i in enumerate(todo):
                if files[j] is not None:
                    logger.info(f"Compressed to {len(files[j]) / 1024 / 1024:.2f}MB at quality={quality}")
                    out[i] = files[j]
                else:
                    keep.append(j)
            if keep and len(keep) < len(todo):
                cur = cur[keep].contiguous()   # a different batch: its DCT is recomputed at the next quality
                fresh = True
            todo = [todo[j] for j in keep]
            quality -= 10
        if todo:
            logger.warning("Quality reduction not enough, also resizing image")
            _, sizes = enc.encode(cur, min_quality, optimize=False, max_bytes=0)   # reference :543: no optimize here
            for j, i in enumerate(todo):
                scale = (target / int(sizes[j])) ** 0.5
                new_w, new_h = int(cur.shape[2] * scale), int(cur.shape[1] * scale)
                small = ops.resize_lanczos(cur[j:j + 1], new_w, new_h)
                out[i] = ops.jpeg_encode(small, min_quality, optimize=True)[0]
                logger.info(f"Compressed to {len(out[i]) / 1024 / 1024:.2f}MB after resize to {(new_w, new_h)}")
        return out  # type: ignore[return-value]

    def _jpeg_decoder(self) -> "ops.JpegDecoder":
        d = getattr(self._tls, "jpeg_dec", None)   # workspace + pinned staging are per calling thread
        if d is None:
            d = self._tls.jpeg_dec = ops.JpegDecoder()
        return d

    def _jpeg_encoder(self) -> "ops.JpegEncoder":
        enc = getattr(self._tls, "jpeg", None)     # workspace + cached DCT: one per calling thread
        if enc is None:
            enc = self._tls.jpeg = ops.JpegEncoder()
        return enc

    def compress_for_azure(self, image: Image.Image, target_size_mb: float = 2.0, initial_quality: int = 95,
                           min_quality: int = 30) -> bytes:
        """Reference :496-557 for one PIL image; the JPEG encoding itself runs on the GPU
        and returns the bytes Pillow's encoder would."""
        if image.mode in ("RGBA", "P", "L", "RGBX"):   # RGBX: Pillow's JPEG writer reads it as RGB (pad dropped)
            image = image.convert("RGB")
        if image.mode != "RGB":
            raise OSError(f"cannot write mode {image.mode} as JPEG")   # Pillow's own error for unsupported modes
        return self.compress_pages_for_azure(self._to_device(image), target_size_mb, initial_quality, min_quality)[0]

    def preprocess_device(self, x: torch.Tensor, apply_deskew: bool = True, apply_binarize: bool = False,
                          apply_contrast: bool = True, apply_sharpness: bool = True):
        """The pixel part of preprocess_for_azure on a resident page batch [N,H,W,C]:
        resize -> deskew -> (contrast 1.2 -> sharpness 1.1 | adaptive binarize).  Returns (pages, angles)."""
        x = ops.resize_if_needed(x, self.max_dimension)
        angles = np.zeros(x.shape[0], np.float64)
        if apply_deskew:
            x, angles = ops.deskew(x)
        if apply_binarize:
            x = ops.adaptive_binarize(x, 2, self.cv_dispatch)
        elif apply_contrast and apply_sharpness:
            x = ops.contrast_sharpness(x, 1.2, 1.1)
        elif apply_contrast:
            x = ops.enhance_contrast(x, 1.2)
        elif apply_sharpness:
            x = ops.enhance_sharpness(x, 1.1)
        return x, angles

    def preprocess_for_azure(self, image: ImageSource, apply_deskew: bool = True, apply_binarize: bool = False,
                             apply_contrast: bool = True, apply_sharpness: bool = True,
                             target_size_mb: float = 2.0) -> bytes:
        img = self.auto_orient(self._open(image))
        logger.info(f"Preprocessing image: {img.size}, mode={img.mode}")
        if img.mode not in ("RGB", "L"):   # any other PIL object: the reference's steps, one drop-in method at a time
            img = self.resize_if_needed(img)
            if apply_deskew:
                img, _ = self.deskew(img)
            if apply_binarize:
                img = self.adaptive_binarize(img)
            else:
                if apply_contrast:
                    img = self.enhance_contrast(img, factor=1.2)
                if apply_sharpness:
                    img = self.enhance_sharpness(img, factor=1.1)
            return self.compress_for_azure(img, target_size_mb=target_size_mb)
        x, _ = self.preprocess_device(self._to_device(img), apply_deskew, apply_binarize, apply_contrast,
                                      apply_sharpness)
        return self.compress_pages_for_azure(x, target_size_mb=target_size_mb)[0]


    def preprocess_pages_for_azure(self, images: Sequence[ImageSource], apply_deskew: bool = True, apply_binarize: bool = False,
                                   apply_contrast: bool = True, apply_sharpness: bool = True,
                                   target_size_mb: float = 2.0) -> List[bytes]:
        """``preprocess_for_azure`` for many pages at once (the page loop of
        ``OCRService.process_pdf_as_images_sync``, ocr_service.py:613-624, as one batched submission).
        Pages are grouped by (size, mode) -- the pages of one PDF share both -- and every group goes through the
        device chain and the JPEG ladder as one batch; results come back in input order and are the same bytes the
        per-page call returns."""
        out: List[Optional[bytes]] = [None] * len(images)
        # Encoded inputs (bytes / paths): baseline JPEG files are decoded in HBM (ops.JpegDecoder: the raster equals
        # Pillow's, so the result is the same bytes) -- the file crosses PCIe instead of the raster and no host core
        # runs libjpeg.  Pillow still parses the header (lazy open: no pixel is decoded) for the EXIF orientation;
        # tagged images, other formats and JPEG flavours outside the device subset take the host codec below.
        imgs: List[Optional[Image.Image]] = [None] * len(images)
        enc_groups = {}
        host_idx: List[int] = []
        for i, src in enumerate(images):
            data = None
            if isinstance(src, bytes):
                data = src
            elif isinstance(src, (str, Path)) and str(src).lower().endswith((".jpg", ".jpeg")) and Path(src).exists():
                data = Path(src).read_bytes()
            info = ops.jpeg_probe(data) if data is not None and data[:2] == b"\xff\xd8" else None
            if info is not None and Image.open(io.BytesIO(data)).getexif().get(0x0112) not in (2, 3, 4, 5, 6, 7, 8):
                enc_groups.setdefault((info.width, info.height, info.channels, info.hs, info.vs), []).append((i, data, info))
            else:
                host_idx.append(i)
        # everything else (PNG -- what pdf_to_images asks poppler for --, TIFF, progressive / CMYK JPEG, PIL objects):
        # the host codec, as in the reference; encoded inputs are decoded on a thread pool (Pillow's decoders release
        # the GIL), not one page after the other
        def _decode(i):
            im = self._open(images[i])
            im.load()
            return im

        enc_host = [i for i in host_idx if not isinstance(images[i], Image.Image)]
        if len(enc_host) > 1:
            with ThreadPoolExecutor(max_workers=min(8, len(enc_host))) as ex:
                for i, im in zip(enc_host, ex.map(_decode, enc_host)):
                    imgs[i] = im
        for i in host_idx:
            if imgs[i] is None:
                imgs[i] = self._open(images[i])
        cap = self._MAX_BATCH_PAGES
        # a group larger than the cap (a 1000-page PDF) goes through in slices: pinned staging and HBM stay bounded
        # (64 A4 pages = 2.2 GB of pinned RGBX + 1.7 GB of rasters); pages are independent, so the bytes do not change
        enc_chunks = [m[k:k + cap] for m in enc_groups.values() for k in range(0, len(m), cap)]
        for members in enc_chunks:
            dec = self._jpeg_decoder()
            blob, offs = dec.pack([d for _, d, _ in members])
            x, status = dec.decode(blob, offs, self.device, members[0][2])
            bad = status.cpu().numpy()            # synchronises: the packed blob may be reused afterwards
            ok = [j for j in range(len(members)) if bad[j] == 0]
            for j in range(len(members)):
                if bad[j] != 0:                   # truncated / corrupt entropy data: Pillow's tolerant decoder takes it
                    imgs[members[j][0]] = self._open(members[j][1])
            if ok:
                xs = x if len(ok) == len(members) else x[torch.as_tensor(ok, device=x.device)]
                xs, _ = self.preprocess_device(xs, apply_deskew, apply_binarize, apply_contrast, apply_sharpness)
                for j, b in zip(ok, self.compress_pages_for_azure(xs, target_size_mb=target_size_mb)):
                    out[members[j][0]] = b
        # EXIF orientation (reference :171-173): pages without the tag -- every rasterised PDF page -- are used as
        # they are (exif_transpose would only copy them); the rare tagged image takes the per-image GPU transpose
        imgs = [im if im is None or im.getexif().get(0x0112) not in (2, 3, 4, 5, 6, 7, 8) else self.auto_orient(im) for im in imgs]
        groups = {}
        for i, im in enumerate(imgs):
            if im is not None:
                groups.setdefault((im.size, im.mode), []).append(i)
        for ((w, h), mode), members in groups.items():
            if mode not in ("RGB", "L"):   # rare: such objects take the per-image path
                for i in members:
                    out[i] = self.preprocess_for_azure(imgs[i], apply_deskew, apply_binarize, apply_contrast,
                                                       apply_sharpness, target_size_mb)
                continue
            for k in range(0, len(members), cap):
                idx = members[k:k + cap]
                x = self._upload([imgs[i] for i in idx])
                x, _ = self.preprocess_device(x, apply_deskew, apply_binarize, apply_contrast, apply_sharpness)
                for i, b in zip(idx, self.compress_pages_for_azure(x, target_size_mb=target_size_mb)):
                    out[i] = b
        return out  # type: ignore[return-value]

    def _host_stage(self, n: int, h: int, w: int, c: int) -> torch.Tensor:
        """Persistent pinned staging buffer for page uploads (grown on demand, reused across calls)."""
        st = getattr(self._tls, "stage", None)     # one per calling thread (see __init__)
        if st is None or st.shape[1:] != (h, w, c) or st.shape[0] < n:
            st = self._tls.stage = torch.empty((n, h, w, c), dtype=torch.uint8, pin_memory=True)
        return st


class _LazySingleton:
    """``image_preprocessor`` singleton (reference :632) built on first use, so importing
    the module never touches CUDA.  ``isinstance(image_preprocessor, ImagePreprocessor)`` holds, as in the reference."""

    _inst: Optional[ImagePreprocessor] = None
    _lock = threading.Lock()

    @property
    def __class__(self):       # isinstance() consults obj.__class__ after type(obj)
        return ImagePreprocessor

    def __getattr__(self, name):
        if _LazySingleton._inst is None:
            with _LazySingleton._lock:
                if _LazySingleton._inst is None:
                    _LazySingleton._inst = ImagePreprocessor()
        return getattr(_LazySingleton._inst, name)


image_preprocessor = _LazySingleton()
