"""Generates tests/golden/dropin_golden.json by running the REAL reference module
(/root/reference/backend/utils/image_preprocessing.py, unmodified, imported in place) through its PUBLIC
methods on PIL objects of every mode, with the flag combinations of optimize_for_ocr / preprocess_for_azure
and the per-page resize loop of pdf_to_images (:287-292).  Run in the build container only.

    python tests/golden/make_dropin_golden.py
"""
import json
import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference/backend")

import cv2  # noqa: E402
import numpy as np  # noqa: E402
import PIL  # noqa: E402
from PIL import Image  # noqa: E402
from utils.image_preprocessing import ImagePreprocessor  # noqa: E402  (the reference itself)

import oracle as O  # noqa: E402  (deterministic synthetic inputs only)
from dropin_images import MAX_DIM, MODES, digest, image_in_mode, mixed_pdf_pages, run  # noqa: E402


METHODS = {
    "resize_if_needed": lambda ip, im: ip.resize_if_needed(im),
    "enhance_contrast": lambda ip, im: ip.enhance_contrast(im, 1.2),
    "enhance_sharpness": lambda ip, im: ip.enhance_sharpness(im, 1.1),
    "denoise": lambda ip, im: ip.denoise(im),
    "convert_to_grayscale": lambda ip, im: ip.convert_to_grayscale(im),
    "binarize": lambda ip, im: ip.binarize(im, 120),
    "adaptive_binarize": lambda ip, im: ip.adaptive_binarize(im),
    "deskew": lambda ip, im: ip.deskew(im),
    "optimize_for_ocr": lambda ip, im: ip.optimize_for_ocr(im),
    "preprocess_for_azure": lambda ip, im: ip.preprocess_for_azure(im),
    "compress_for_azure": lambda ip, im: ip.compress_for_azure(im, target_size_mb=0.05),
}


def main():
    cv2.setUseOptimized(False)
    ip = ImagePreprocessor(max_dimension=MAX_DIM)
    cases = []
    for mode in MODES:
        for name, fn in METHODS.items():
            im = image_in_mode(O, mode, 0)
            cases.append({"kind": "mode", "mode": mode, "method": name, "want": run(lambda: fn(ip, im))})
    flags = [dict(grayscale=True), dict(apply_denoise=True), dict(grayscale=True, apply_denoise=True),
             dict(apply_contrast=False), dict(apply_sharpness=False), dict(apply_contrast=False, apply_sharpness=False)]
    for mode in ("RGB", "L"):
        for seed in (0, 1):
            for kw in flags:
                im = image_in_mode(O, mode, seed)
                cases.append({"kind": "optimize_for_ocr", "mode": mode, "seed": seed, "kwargs": kw,
                              "want": run(lambda: ip.optimize_for_ocr(im, **kw))})
    az = [dict(apply_binarize=True), dict(apply_deskew=False), dict(apply_contrast=False), dict(apply_sharpness=False),
          dict(target_size_mb=0.05)]
    for mode in ("RGB", "L"):
        for kw in az:
            im = image_in_mode(O, mode, 1)
            cases.append({"kind": "preprocess_for_azure", "mode": mode, "seed": 1, "kwargs": kw,
                          "want": run(lambda: ip.preprocess_for_azure(im, **kw))})
    pages = mixed_pdf_pages(O)
    cases.append({"kind": "resize_pages", "want": [digest(ip.resize_if_needed(p)) for p in pages]})
    out = {"generator": "tests/golden/make_dropin_golden.py (runs the unmodified reference module)",
           "versions": {"Pillow": PIL.__version__, "opencv": cv2.__version__, "numpy": np.__version__},
           "max_dimension": MAX_DIM, "cases": cases}
    path = os.path.join(HERE, "dropin_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print(f"{len(cases)} cases -> {path}")


if __name__ == "__main__":
    main()
