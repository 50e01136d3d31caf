"""Generates tests/golden/angle_golden.json: the deskew angle of the REAL reference module
(/root/reference/backend/utils/image_preprocessing.py:372-460, unmodified, imported in place) on synthetic pages,
as float64 hex, together with what glibc's atan2 gives for the same segments.

    python tests/golden/make_angle_golden.py

Why a file of its own: the reference computes np.degrees(np.arctan2(dy, dx)) per Hough segment (:421).  numpy's
arctan2 is NOT glibc's atan2 on AVX-512 builds (numpy dispatches to its bundled SIMD math there): ~0.3 % of the
segments differ in the last place, and when such a segment is the median the page's angle differs (seeds 154, 501, 506 below:
3 of the first 600 synthetic pages).
The product and the oracle therefore take the per-line angle from numpy; these goldens pin that choice and record
the numpy build / CPU dispatch they were produced with.  Run in the build container only."""
import json
import math
import os
import sys

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/backend")

import cv2  # noqa: E402
import numpy as np  # noqa: E402
from PIL import Image  # noqa: E402
from utils.image_preprocessing import ImagePreprocessor  # noqa: E402  (the reference itself)

import oracle as O  # noqa: E402  (deterministic synthetic inputs + the glibc restatement for the record)

SEEDS = list(range(150, 158)) + [501, 506]   # 154, 501, 506: pages whose median segment is one of the differing ones
H, W, MD = 877, 620, 400


def main():
    cv2.setUseOptimized(False)
    ip = ImagePreprocessor(max_dimension=MD)
    cases = []
    for seed in SEEDS:
        pil = ip.resize_if_needed(Image.fromarray(O.synth_page(H, W, seed)))
        _, angle = ip.deskew(pil)
        gray = cv2.cvtColor(cv2.cvtColor(np.array(pil), cv2.COLOR_RGB2BGR), cv2.COLOR_BGR2GRAY)
        lines = cv2.HoughLinesP(cv2.Canny(gray, 50, 150, apertureSize=3), 1, np.pi / 180, threshold=100, minLineLength=100,
                                maxLineGap=10)
        libm = O.median_angle_libm(lines.reshape(-1, 4)) if lines is not None else 0.0
        libm_result = libm if 0.5 <= abs(libm) <= 45 or abs(libm) < 0.5 else 0.0
        cases.append(dict(seed=seed, h=H, w=W, max_dim=MD, angle_hex=float(angle).hex(), angle=float(angle),
                          glibc_median_hex=float(libm_result).hex(), n_lines=0 if lines is None else int(len(lines))))
    try:
        from numpy._core._multiarray_umath import __cpu_features__ as feats
        simd = bool(feats.get("AVX512_SKX"))   # numpy's bundled SVML loops (arctan2 among them) are built for AVX512_SKX
    except Exception:  # noqa: BLE001
        simd = False
    ys, xs = np.meshgrid(np.arange(-40, 41), np.arange(100, 679), indexing="ij")
    a = np.degrees(np.arctan2(ys.ravel(), xs.ravel()))
    b = np.array([math.atan2(int(p), int(q)) for p, q in zip(ys.ravel(), xs.ravel())]) * (180.0 / math.pi)
    out = dict(numpy=np.__version__, opencv=cv2.__version__, numpy_avx512_skx=simd,
               arctan2_vs_glibc_mismatch_fraction=float((a != b).mean()),
               differing=[c["seed"] for c in cases if c["angle_hex"] != c["glibc_median_hex"]], cases=cases)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "angle_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(len(cases), "cases; pages whose angle differs between numpy and glibc:", out["differing"])


if __name__ == "__main__":
    main()
