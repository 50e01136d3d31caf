"""Deterministic PIL inputs for the drop-in surface goldens (shared by the generator, which runs the
unmodified reference, and by tests/test_gpu_dropin.py, which runs the CUDA drop-in)."""
import numpy as np
from PIL import Image

H, W, MAX_DIM = 877, 620, 600          # A4 @ 75 dpi; max_dimension=600 forces the Lanczos resize
MODES = ["RGBA", "LA", "RGBX", "CMYK", "YCbCr", "HSV", "P", "1"]


def base_rgb(O, seed: int) -> np.ndarray:
    return O.synth_page(H, W, seed)


def image_in_mode(O, mode: str, seed: int = 0) -> Image.Image:
    rgb = base_rgb(O, seed)
    rng = np.random.default_rng(1000 + seed)
    pil = Image.fromarray(rgb)
    if mode == "RGB":
        return pil
    if mode == "L":
        return pil.convert("L")
    if mode == "RGBA":
        a = rng.integers(0, 256, size=(H, W, 1), dtype=np.uint8)
        return Image.fromarray(np.concatenate([rgb, a], axis=-1))
    if mode == "LA":
        a = rng.integers(0, 256, size=(H, W), dtype=np.uint8)
        return Image.fromarray(np.stack([np.asarray(pil.convert("L")), a], axis=-1))
    if mode == "RGBX":
        x = rng.integers(0, 256, size=(H, W, 1), dtype=np.uint8)
        return Image.frombytes("RGBX", (W, H), np.concatenate([rgb, x], axis=-1).tobytes())
    if mode == "P":
        return pil.quantize(64, method=Image.Quantize.MEDIANCUT, dither=Image.Dither.NONE)
    if mode == "1":
        return pil.convert("L").point(lambda v: 255 if v > 140 else 0, "1")
    return pil.convert(mode)            # CMYK, YCbCr, HSV


def mixed_pdf_pages(O):
    """A 'PDF' whose pages differ in size and mode (resize_pages groups them by shape)."""
    pages = []
    for i, (h, w) in enumerate([(877, 620), (620, 877), (877, 620), (400, 300), (877, 620)]):
        a = O.synth_page(h, w, 40 + i)
        im = Image.fromarray(a)
        pages.append(im.convert("L") if i == 2 else im)
    return pages


def digest(out):
    """What the goldens store for a method result: PIL image, (image, angle) or JPEG bytes."""
    import hashlib

    if isinstance(out, tuple):      # deskew -> (image, angle)
        d = digest(out[0])
        d["angle"] = float(out[1])
        return d
    if isinstance(out, (bytes, bytearray)):
        return {"bytes": len(out), "sha": hashlib.sha256(out).hexdigest()}
    img = out
    raw = img.tobytes() if img.mode != "1" else img.convert("L").tobytes()
    return {"mode": img.mode, "size": list(img.size), "sha": hashlib.sha256(raw).hexdigest()}


def run(fn):
    try:
        return digest(fn())
    except Exception as e:  # noqa: BLE001 - the exception type IS the golden
        return {"raises": type(e).__name__}
