"""Generates tests/golden/reference_golden.json by running the REAL reference module
(/root/reference/backend/utils/image_preprocessing.py, unmodified, imported in place) on
deterministic synthetic pages.  Run in the build container only (the reference tree does not
travel to the GPU box); the JSON it writes is committed and is what pins the oracle.

    python tests/golden/make_golden.py

Library versions are recorded: the reference pins none (requirements.txt:20-23)."""
import hashlib
import io
import json
import os
import sys

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/backend")

import cv2  # noqa: E402
import numpy as np  # noqa: E402
import PIL  # noqa: E402
from PIL import Image  # noqa: E402
from utils.image_preprocessing import ImagePreprocessor  # noqa: E402  (the reference itself)

import oracle as O  # noqa: E402  (only for the deterministic synthetic inputs)


def sha(a) -> str:
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def main():
    cv2.setUseOptimized(False)  # dispatch-independent float path for adaptiveThreshold
    cases = []
    pages = {
        "a4_quarter": (877, 620),   # A4 @ 75 dpi
        "a4_third": (1169, 827),    # A4 @ 100 dpi
        "landscape": (620, 877),
    }
    for name, (h, w) in pages.items():
        for seed in (0, 1, 2):
            rgb = O.synth_page(h, w, seed)
            pil = Image.fromarray(rgb)
            gray = pil.convert("L")
            for md in (400, 600):
                ip = ImagePreprocessor(max_dimension=md)
                key = dict(page=name, h=h, w=w, seed=seed, max_dim=md)
                r = ip.resize_if_needed(pil)
                cases.append(dict(key, stage="resize_if_needed", size=list(r.size), sha=sha(np.asarray(r))))
                rg = ip.resize_if_needed(gray)
                cases.append(dict(key, stage="resize_if_needed_L", size=list(rg.size), sha=sha(np.asarray(rg))))
                d, angle = ip.deskew(r)
                rn = np.asarray(r)
                g = cv2.cvtColor(cv2.cvtColor(rn, cv2.COLOR_RGB2BGR), cv2.COLOR_BGR2GRAY)
                e = cv2.Canny(g, 50, 150, apertureSize=3)
                ln = cv2.HoughLinesP(e, 1, np.pi / 180, threshold=100, minLineLength=100, maxLineGap=10)
                ln = np.zeros((0, 4), np.int32) if ln is None else ln[:, 0, :]
                cases.append(dict(key, stage="deskew", angle=angle, sha=sha(np.asarray(d)), gray_cv_sha=sha(g),
                                  canny_sha=sha(e), n_edges=int((e > 0).sum()), n_lines=int(len(ln)),
                                  lines_sha=sha(ln.astype(np.int32))))
                chain = ip.enhance_sharpness(ip.enhance_contrast(d, 1.2), 1.1)
                cases.append(dict(key, stage="azure_chain_before_jpeg", sha=sha(np.asarray(chain))))
                ob = ip.optimize_for_ocr(pil)
                cases.append(dict(key, stage="optimize_for_ocr", sha=sha(np.asarray(ob))))
                ab = ip.adaptive_binarize(r)
                cases.append(dict(key, stage="adaptive_binarize", sha=sha(np.asarray(ab))))
            ip = ImagePreprocessor(max_dimension=2000)
            key = dict(page=name, h=h, w=w, seed=seed)
            cases.append(dict(key, stage="convert_to_grayscale", sha=sha(np.asarray(ip.convert_to_grayscale(pil)))))
            for f in (1.2, 1.3):
                cases.append(dict(key, stage="enhance_contrast", factor=f, sha=sha(np.asarray(ip.enhance_contrast(pil, f)))))
                cases.append(dict(key, stage="enhance_contrast_L", factor=f, sha=sha(np.asarray(ip.enhance_contrast(gray, f)))))
            for f in (1.1, 1.2):
                cases.append(dict(key, stage="enhance_sharpness", factor=f, sha=sha(np.asarray(ip.enhance_sharpness(pil, f)))))
            cases.append(dict(key, stage="denoise", sha=sha(np.asarray(ip.denoise(pil)))))
            b = ip.binarize(pil)
            cases.append(dict(key, stage="binarize", mode=b.mode, sha=sha(np.asarray(b.convert("L")))))
            if seed == 0:
                for o in range(1, 9):
                    ex = pil.getexif()
                    ex[0x0112] = o
                    buf = io.BytesIO()
                    pil.save(buf, format="PNG", exif=ex)
                    q = Image.open(io.BytesIO(buf.getvalue()))
                    t = ip.auto_orient(q)
                    cases.append(dict(key, stage="auto_orient", orientation=o, size=list(t.size), sha=sha(np.asarray(t))))
    # one full-size A4 page through the default resize + deskew (the headline geometry)
    rgb = O.synth_page(3508, 2480, 0)
    for md in (960, 2000):
        ip = ImagePreprocessor(max_dimension=md)
        r = ip.resize_if_needed(Image.fromarray(rgb))
        d, angle = ip.deskew(r)
        cases.append(dict(page="a4_300dpi", h=3508, w=2480, seed=0, max_dim=md, stage="resize_if_needed",
                          size=list(r.size), sha=sha(np.asarray(r))))
        cases.append(dict(page="a4_300dpi", h=3508, w=2480, seed=0, max_dim=md, stage="deskew_full", angle=angle,
                          sha=sha(np.asarray(d))))
    out = {
        "generator": "tests/golden/make_golden.py (runs the unmodified reference module)",
        "reference_module": "backend/utils/image_preprocessing.py",
        "versions": {"Pillow": PIL.__version__, "opencv": cv2.__version__, "numpy": np.__version__},
        "cv2_setUseOptimized": False,
        "cases": cases,
    }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print(f"{len(cases)} cases -> {path}")


if __name__ == "__main__":
    main()
