"""Randomised sweep of the drop-in surface (ocr-system_b200/image_preprocessing.py on the GPU) against the REAL
reference module running beside it on the host (oracle/_ref/image_preprocessing.py: the unmodified copy that
oracle/make_ref.py places and the snapshot carries to the GPU box): random PIL images of random modes, sizes and
contents through every public method with random arguments.  Results must be the same object for object: PIL mode,
size and pixel bytes; float64 angle; JPEG file bytes; or the same exception type.

    python tools/sweep_dropin_vs_reference.py --seeds 0 200
"""
import argparse
import io
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from PIL import Image

import oracle as O
from oracle import reference_port as RP

MODES = ["RGB", "RGB", "RGB", "L", "L", "RGBA", "LA", "P", "1", "CMYK", "YCbCr", "RGBX"]


TINY = False   # --tiny: degenerate sizes (1 x 1 ... 24 x 24) and extreme aspect ratios


def make_image(rng, seed):
    h, w = int(rng.integers(24, 1400)), int(rng.integers(24, 1400))
    if TINY:
        pick = seed % 4
        if pick == 0:
            h, w = int(rng.integers(1, 25)), int(rng.integers(1, 25))
        elif pick == 1:
            h, w = int(rng.integers(1, 6)), int(rng.integers(200, 3000))
        elif pick == 2:
            h, w = int(rng.integers(200, 3000)), int(rng.integers(1, 6))
        else:
            h, w = int(rng.integers(8, 40)), int(rng.integers(8, 40))
    kind = seed % 4
    if kind == 0:
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    elif kind == 1:
        yy, xx = np.mgrid[0:h, 0:w]
        rgb = np.stack([(xx * 255 // w), (yy * 255 // h), ((xx + yy) * 255 // (w + h))], -1).astype(np.uint8)
        rgb[h // 3: h // 3 + 2, w // 8: 7 * w // 8] = 15
    else:
        rgb = O.synth_page(h, w, seed)
    pil = Image.fromarray(rgb)
    mode = MODES[int(rng.integers(0, len(MODES)))]
    if mode == "RGB":
        return pil
    if mode == "RGBA":
        a = rng.integers(0, 256, size=(h, w, 1), dtype=np.uint8)
        return Image.fromarray(np.concatenate([rgb, a], axis=-1))
    if mode == "LA":
        a = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
        return Image.fromarray(np.stack([np.asarray(pil.convert("L")), a], axis=-1))
    if mode == "RGBX":
        x = rng.integers(0, 256, size=(h, w, 1), dtype=np.uint8)
        return Image.frombytes("RGBX", (w, h), np.concatenate([rgb, x], axis=-1).tobytes())
    if mode == "P":
        return pil.quantize(64, method=Image.Quantize.MEDIANCUT, dither=Image.Dither.NONE)
    if mode == "1":
        return pil.convert("L").point(lambda v: 255 if v > 140 else 0, "1")
    return pil.convert(mode)


def digest(out):
    if isinstance(out, tuple):
        d = digest(out[0])
        d["angle"] = float(out[1]).hex()
        return d
    if isinstance(out, (bytes, bytearray)):
        return {"bytes": bytes(out)}
    raw = out.tobytes() if out.mode != "1" else out.convert("L").tobytes()
    return {"mode": out.mode, "size": tuple(out.size), "raw": raw}


def run(fn):
    try:
        return digest(fn())
    except Exception as e:  # noqa: BLE001 - the exception type is part of the contract
        return {"raises": type(e).__name__}


def sweep(seed_lo: int, seed_hi: int, verbose: bool = True):
    """Returns (calls compared, list of mismatches), or None when oracle/_ref is not in this snapshot."""
    ref_mod = RP.real_preprocessor()
    if ref_mod is None:
        return None
    from ocr_system_b200.image_preprocessing import ImagePreprocessor

    # documented deviations (DESIGN "PIL modes"): the device JPEG encoder writes 3-component YCbCr files only
    jpeg_deviation = {"CMYK", "YCbCr", "1"}
    bad, checked = [], 0
    for seed in range(seed_lo, seed_hi):
        rng = np.random.default_rng(seed)
        img = make_image(rng, seed)
        md = int(rng.choice([128, 300, 600, 960, 2000] + ([4, 16, 33] if TINY else [])))
        ref = ref_mod.ImagePreprocessor(max_dimension=md)
        # the live cv2 beside us runs its default dispatch: tell the drop-in to reproduce that one (its default)
        ours = ImagePreprocessor(max_dimension=md)
        cf, sf, thr = float(rng.choice([1.0, 1.2, 1.3, 0.8])), float(rng.choice([1.0, 1.1, 1.5, 0.5])), int(rng.integers(0, 256))
        flags = dict(auto_resize=bool(rng.integers(0, 2)), enhance_contrast=bool(rng.integers(0, 2)),
                     enhance_sharpness=bool(rng.integers(0, 2)), grayscale=bool(rng.integers(0, 2)),
                     apply_denoise=bool(rng.integers(0, 2)))
        az = dict(apply_deskew=bool(rng.integers(0, 2)), apply_binarize=bool(rng.integers(0, 2)),
                  target_size_mb=float(rng.choice([2.0, 0.2, 0.05, 0.01])))
        calls = {
            "resize_if_needed": lambda ip: ip.resize_if_needed(img),
            "enhance_contrast": lambda ip: ip.enhance_contrast(img, cf),
            "enhance_sharpness": lambda ip: ip.enhance_sharpness(img, sf),
            "denoise": lambda ip: ip.denoise(img),
            "convert_to_grayscale": lambda ip: ip.convert_to_grayscale(img),
            "binarize": lambda ip: ip.binarize(img, thr),
            "adaptive_binarize": lambda ip: ip.adaptive_binarize(img),
            "deskew": lambda ip: ip.deskew(img),
            "optimize_for_ocr": lambda ip: ip.optimize_for_ocr(img, **flags),
            "preprocess_for_azure": lambda ip: ip.preprocess_for_azure(img, **az),
            "compress_for_azure": lambda ip: ip.compress_for_azure(img, target_size_mb=az["target_size_mb"]),
        }
        for name, fn in calls.items():
            want, got = run(lambda: fn(ref)), run(lambda: fn(ours))
            checked += 1
            if want != got:
                if name in ("compress_for_azure", "preprocess_for_azure") and img.mode in jpeg_deviation and "raises" in got:
                    continue
                info = dict(seed=seed, method=name, mode=img.mode, size=img.size, md=md,
                            want={k: (v if k not in ("raw", "bytes") else len(v)) for k, v in want.items()},
                            got={k: (v if k not in ("raw", "bytes") else len(v)) for k, v in got.items()})
                if name == "optimize_for_ocr":
                    info["flags"] = flags
                if name in ("preprocess_for_azure", "compress_for_azure"):
                    info["az"] = az
                if "raw" in want and "raw" in got and len(want["raw"]) == len(got["raw"]):
                    info["differing_bytes"] = int((np.frombuffer(want["raw"], np.uint8) != np.frombuffer(got["raw"], np.uint8)).sum())
                bad.append(info)
                if verbose:
                    print("MISMATCH", json.dumps(info, default=str), flush=True)
    return checked, bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, nargs=2, default=[0, 100])
    ap.add_argument("--tiny", action="store_true")
    a = ap.parse_args()
    import cv2

    global TINY
    TINY = a.tiny
    t0 = time.time()
    res = sweep(*a.seeds)
    if res is None:
        print(json.dumps({"unavailable": "oracle/_ref/image_preprocessing.py is not in this snapshot"}))
        return
    checked, bad = res
    print(json.dumps({"seeds": a.seeds, "checked": checked, "mismatches": len(bad), "seconds": round(time.time() - t0, 1),
                      "cv2_optimized": bool(cv2.useOptimized())}))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/sweep_dropin_vs_reference.json", "w") as f:
        json.dump({"seeds": a.seeds, "checked": checked, "mismatches": bad}, f, default=str)


if __name__ == "__main__":
    main()
