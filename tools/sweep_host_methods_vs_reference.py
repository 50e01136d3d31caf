"""Host-only methods of the drop-in ImagePreprocessor (no device work: load_image, load_image_bytes, get_image_info,
image_to_bytes, save_image, get_optimal_size) against the UNMODIFIED reference module (/root/reference, build container
only) on random images of every mode / container format Pillow writes here.  Returned objects (mode, size, pixel bytes),
dictionaries, file bytes and exception types must be identical.

    python tools/sweep_host_methods_vs_reference.py
"""
import io
import logging
import os
import sys
import tempfile

import numpy as np
from PIL import Image

logging.disable(logging.CRITICAL)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference/backend")
sys.path.insert(0, ROOT)
from utils.image_preprocessing import ImagePreprocessor as Ref  # noqa: E402

from ocr_system_b200.image_preprocessing import ImagePreprocessor as Ours  # noqa: E402

MODES = ["RGB", "L", "RGBA", "P", "1", "CMYK", "LA", "I;16", "F", "I"]
FORMATS = {"PNG": ["RGB", "L", "RGBA", "P", "1", "LA", "I;16", "I"], "JPEG": ["RGB", "L", "CMYK"], "TIFF": MODES, "BMP": ["RGB", "L", "P", "1"],
           "WEBP": ["RGB", "RGBA"], "GIF": ["L", "P"]}


def rand_image(rng, mode):
    h, w = int(rng.integers(1, 60)), int(rng.integers(1, 60))
    rgb = Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
    if mode == "I;16":
        return Image.fromarray(rng.integers(0, 65536, (h, w)).astype(np.uint16))
    if mode == "F":
        return Image.fromarray(rng.random((h, w)).astype(np.float32) * 255)
    if mode == "I":
        return Image.fromarray(rng.integers(-1000, 70000, (h, w)).astype(np.int32))
    return rgb.convert(mode)


def outcome(fn):
    try:
        r = fn()
    except Exception as e:  # noqa: BLE001
        return ("raises", type(e).__name__)
    if isinstance(r, Image.Image):
        return ("image", r.mode, r.size, r.tobytes())
    if isinstance(r, (str, os.PathLike)):
        return ("path", os.path.basename(str(r)), open(r, "rb").read())
    return ("value", r)


def main():
    rng = np.random.default_rng(0)
    ref, ours = Ref(max_dimension=int(rng.integers(20, 80))), None
    ours = Ours(max_dimension=ref.max_dimension)
    n = bad = 0
    with tempfile.TemporaryDirectory() as tmp:
        for it in range(600):
            fmt = list(FORMATS)[it % len(FORMATS)]
            mode = FORMATS[fmt][int(rng.integers(0, len(FORMATS[fmt])))]
            im = rand_image(rng, mode)
            buf = io.BytesIO()
            try:
                im.save(buf, format=fmt)
            except Exception:  # noqa: BLE001 - Pillow cannot write this combination
                continue
            data = buf.getvalue()
            path = os.path.join(tmp, f"in_{it}.{fmt.lower()}")
            open(path, "wb").write(data)
            q = int(rng.integers(1, 100))
            out_fmt = ["PNG", "JPEG", "jpg", "TIFF", "BMP"][int(rng.integers(0, 5))]
            suffix = [".jpg", ".jpeg", ".png", ".tif", ".bmp", ".PNG"][int(rng.integers(0, 6))]
            checks = {
                "load_image": lambda p: p.load_image(path),
                "load_image_missing": lambda p: p.load_image(os.path.join(tmp, "nope.png")),
                "load_image_bytes": lambda p: p.load_image_bytes(data),
                "load_image_bytes_garbage": lambda p: p.load_image_bytes(data[: len(data) // 3][::-1]),
                "get_image_info_path": lambda p: sorted(p.get_image_info(path).items()),
                "get_image_info_image": lambda p: sorted(p.get_image_info(im).items()),
                "image_to_bytes": lambda p: p.image_to_bytes(im, format=out_fmt, quality=q),
                "save_image": lambda p: p.save_image(im, os.path.join(tmp, "out", f"o_{it}{suffix}"), quality=q, optimize=bool(it & 1)),
                "get_optimal_size": lambda p: p.get_optimal_size(im.width * 7, im.height * 5),
            }
            for name, fn in checks.items():
                a, b = outcome(lambda: fn(ref)), outcome(lambda: fn(ours))
                n += 1
                if a != b:
                    bad += 1
                    print("MISMATCH", name, fmt, mode, im.size, a[:3], b[:3])
    print(f"comparisons {n} mismatches {bad}")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
