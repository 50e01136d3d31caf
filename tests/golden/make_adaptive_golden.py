"""Generates tests/golden/adaptive_dispatch_golden.json: what the REAL reference's adaptive_binarize
(/root/reference/backend/utils/image_preprocessing.py:462-494, unmodified) returns under OpenCV's DEFAULT dispatch
(cv2.setUseOptimized(True): what the application runs) and under the plain path, on gray planes where the two differ.

    python tests/golden/make_adaptive_golden.py

cv2.adaptiveThreshold blurs in float32; the AVX2 build of OpenCV's separable filter uses fused multiply-add in its
vector loops, the plain build does not, so ~10 % of pages differ by a pixel between the modes.  The 217 goldens of
make_golden.py use the plain path; these pin the other one (oracle cv_dispatch="avx2", lumina LUMINA_CV_AVX2).
Run in the build container only."""
import hashlib
import json
import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference/backend")

import cv2  # noqa: E402
import numpy as np  # noqa: E402
from PIL import Image  # noqa: E402
from utils.image_preprocessing import ImagePreprocessor  # noqa: E402  (the reference itself)

from adaptive_inputs import plane  # noqa: E402


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ip = ImagePreprocessor(max_dimension=100000)
    cases, differing, same = [], 0, 0
    for seed in range(400):
        g = plane(seed)
        pil = Image.fromarray(g)
        cv2.setUseOptimized(True)
        opt = np.asarray(ip.adaptive_binarize(pil))
        cv2.setUseOptimized(False)
        plain = np.asarray(ip.adaptive_binarize(pil))
        cv2.setUseOptimized(True)
        d = int((opt != plain).sum())
        if (d and differing < 18) or (not d and same < 4):
            differing += bool(d)
            same += not d
            cases.append(dict(seed=seed, h=int(g.shape[0]), w=int(g.shape[1]), w_mod_8=int(g.shape[1] % 8), differing_px=d,
                              sha_default=sha(opt), sha_plain=sha(plain)))
        if differing >= 18 and same >= 4:
            break
    out = dict(generator="tests/golden/make_adaptive_golden.py (runs the unmodified reference module)",
               opencv=cv2.__version__, numpy=np.__version__,
               cpu_features=cv2.getCPUFeaturesLine(),
               avx2=bool(cv2.checkHardwareSupport(10)), fma3=bool(cv2.checkHardwareSupport(12)),   # CV_CPU_AVX2 / CV_CPU_FMA3
               cases=cases)
    with open(os.path.join(HERE, "adaptive_dispatch_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(len(cases), "cases,", differing, "with differing modes; width residues:", sorted({c["w_mod_8"] for c in cases}))


if __name__ == "__main__":
    main()
