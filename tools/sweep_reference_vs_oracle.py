"""Randomised parity sweep: the UNMODIFIED reference module (/root/reference, build container only) against the
oracle on many more inputs than the committed goldens hold -- synthetic pages of random shapes / max_dimensions,
uniform-noise images and smooth gradients.  Every stage of tests/golden/make_golden.py is compared; mismatches are
printed with the seed so that they can be turned into goldens.  This is how the numpy-vs-glibc arctan2 difference of
the deskew angle was found.

    python tools/sweep_reference_vs_oracle.py --seeds 0 200 [--procs 8]
"""
import argparse
import logging
import multiprocessing as mp
import os
import sys

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/backend")


A4 = False


def one(seed: int):
    import cv2
    import numpy as np
    from PIL import Image
    from utils.image_preprocessing import ImagePreprocessor

    import oracle as O

    logging.disable(logging.CRITICAL)
    cv2.setUseOptimized(bool(seed & 1))          # both OpenCV dispatch modes
    rng = np.random.default_rng(seed)
    h, w = int(rng.integers(300, 1300)), int(rng.integers(300, 1300))
    md = int(rng.choice([256, 400, 600, 960, 2000]))
    kind = seed % 4
    if A4:                                        # the headline geometry: A4 at 300 dpi through max_dimension 960 / 2000
        h, w, md, kind = 3508, 2480, (960 if seed & 2 else 2000), 0
    if kind == 3:
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    elif kind == 2:
        yy, xx = np.mgrid[0:h, 0:w]
        rgb = np.stack([(xx * 255 // w), (yy * 255 // h), ((xx + yy) * 255 // (w + h))], -1).astype(np.uint8)
        rgb[h // 3: h // 3 + 3, w // 8: 7 * w // 8] = 10       # a long dark rule: gives HoughLinesP something
    else:
        rgb = O.synth_page(h, w, seed)
    pil = Image.fromarray(rgb)
    ip = ImagePreprocessor(max_dimension=md)
    bad = []

    def cmp(name, ref, got):
        if not np.array_equal(np.asarray(ref), got):
            bad.append((seed, kind, h, w, md, name, int((np.asarray(ref) != got).sum())))

    r = ip.resize_if_needed(pil)
    tw, th = O.target_size(w, h, md)
    ro = rgb if (tw, th) == (w, h) else O.resize_lanczos(rgb, tw, th)
    cmp("resize", r, ro)
    d, angle = ip.deskew(r)
    do, ao, _ = O.deskew(ro)
    if angle != ao:
        bad.append((seed, kind, h, w, md, "angle", repr(angle), repr(ao)))
    cmp("deskew", d, do)
    cmp("contrast+sharpness", ip.enhance_sharpness(ip.enhance_contrast(d, 1.2), 1.1), O.sharpness(O.contrast(do, 1.2), 1.1))
    cmp("adaptive", ip.adaptive_binarize(r).convert("L"), O.adaptive_gauss11(O.gray_pil(ro), 2, cv_dispatch="avx2" if seed & 1 else "plain"))
    cmp("gray", ip.convert_to_grayscale(r), O.gray_pil(ro))
    cmp("denoise", ip.denoise(r), O.median3(ro))
    cmp("binarize", ip.binarize(r).convert("L"), O.threshold(O.gray_pil(ro), 128))
    return bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, nargs=2, default=[0, 64])
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--a4", action="store_true", help="full-size A4 300-dpi pages through max_dimension 960 / 2000")
    a = ap.parse_args()
    global A4
    A4 = a.a4
    with mp.get_context("fork").Pool(a.procs) as pool:
        n_bad = 0
        for i, bad in enumerate(pool.imap_unordered(one, range(*a.seeds), chunksize=1)):
            for b in bad:
                n_bad += 1
                print("MISMATCH", b, flush=True)
    print(f"seeds {a.seeds[0]}..{a.seeds[1] - 1}: {n_bad} mismatching stages", flush=True)


if __name__ == "__main__":
    main()
