"""Randomised sweep of the ENCODED-input path (bytes / files, image_preprocessing.py:57-75, 592-597) against the REAL
reference module beside it (oracle/_ref): random pages saved in the formats the application accepts -- baseline /
progressive / grayscale / CMYK JPEG with random quality, subsampling, restart intervals and EXIF orientation, PNG of
several modes, TIFF (LZW), BMP, GIF, WebP -- handed as bytes to preprocess_for_azure (per file) and to
preprocess_pages_for_azure (as one batch, where baseline JPEGs are decoded on the device).  The JPEG bytes returned must
equal the reference's, file for file.

    python tools/sweep_files_vs_reference.py --seeds 0 60
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


def encode(rng, rgb):
    pil = Image.fromarray(rgb)
    kind = int(rng.integers(0, 12))
    b = io.BytesIO()
    desc = ""
    if kind <= 4:      # baseline JPEG flavours (the device-decode subset and its edges)
        kw = dict(quality=int(rng.choice([30, 60, 75, 90, 95, 100])), optimize=bool(rng.integers(0, 2)),
                  subsampling=int(rng.choice([0, 1, 2])))
        if rng.integers(0, 3) == 0:
            kw["restart_marker_blocks"] = int(rng.choice([1, 3, 8, 50]))
        if rng.integers(0, 4) == 0:
            ex = pil.getexif()
            ex[0x0112] = int(rng.integers(1, 9))
            kw["exif"] = ex
        for attempt in range(3):   # Pillow's own output buffer can be too small for incompressible data
            try:
                b = io.BytesIO()
                pil.save(b, "JPEG", **kw)
                break
            except OSError:
                if attempt == 0:
                    kw["optimize"] = True
                else:
                    kw.pop("restart_marker_blocks", None)
                    kw["quality"] = 90
        if b.getbuffer().nbytes == 0:
            b = io.BytesIO()
            pil.save(b, "PNG")
            kw = {"quality": "png-fallback"}
        desc = f"jpeg {kw.get('quality')} sub{kw.get('subsampling')} rst{kw.get('restart_marker_blocks')} exif{'exif' in kw}"
    elif kind == 5:
        pil.save(b, "JPEG", quality=80, progressive=True); desc = "jpeg progressive"
    elif kind == 6:
        pil.convert("L").save(b, "JPEG", quality=85); desc = "jpeg gray"
    elif kind == 7:
        pil.convert("CMYK").save(b, "JPEG", quality=85); desc = "jpeg cmyk"
    elif kind == 8:
        m = str(rng.choice(["RGB", "L", "P", "RGBA", "1"]))
        (pil.quantize(32) if m == "P" else pil.convert(m)).save(b, "PNG"); desc = "png " + m
    elif kind == 9:
        pil.save(b, "TIFF", compression="tiff_lzw"); desc = "tiff lzw"
    elif kind == 10:
        pil.save(b, "BMP"); desc = "bmp"
    else:
        try:
            pil.save(b, "WEBP", lossless=True); desc = "webp"
        except Exception:  # noqa: BLE001
            b = io.BytesIO(); pil.save(b, "GIF"); desc = "gif"
    return b.getvalue(), desc


def sweep(seed_lo, seed_hi, verbose=True):
    ref_mod = RP.real_preprocessor()
    if ref_mod is None:
        return None
    from ocr_system_b200.image_preprocessing import ImagePreprocessor

    bad, checked = [], 0
    for seed in range(seed_lo, seed_hi):
        rng = np.random.default_rng(seed)
        md = int(rng.choice([300, 600, 960]))
        ref, ours = ref_mod.ImagePreprocessor(max_dimension=md), ImagePreprocessor(max_dimension=md)
        h, w = int(rng.integers(60, 1300)), int(rng.integers(60, 1300))
        files, descs = [], []
        for k in range(int(rng.integers(2, 7))):
            same = k > 0 and rng.integers(0, 2) == 0          # pages of one document share their size
            hh, ww = (h, w) if same else (int(rng.integers(60, 1300)), int(rng.integers(60, 1300)))
            rgb = O.synth_page(hh, ww, seed * 16 + k) if (seed + k) % 3 else rng.integers(0, 256, (hh, ww, 3), dtype=np.uint8)
            f, d = encode(rng, rgb)
            files.append(f); descs.append(f"{d} {hh}x{ww}")
        az = dict(apply_deskew=bool(rng.integers(0, 2)), apply_binarize=bool(rng.integers(0, 2)),
                  target_size_mb=float(rng.choice([2.0, 0.2, 0.03])))

        def call(ip, f):
            try:
                return ip.preprocess_for_azure(f, **az)
            except Exception as e:  # noqa: BLE001
                return "raises " + type(e).__name__

        want = [call(ref, f) for f in files]
        try:
            batch = ours.preprocess_pages_for_azure(files, **az)
        except Exception as e:  # noqa: BLE001
            batch = ["batch raises " + type(e).__name__] * len(files)
        single = [call(ours, f) for f in files]
        for k in range(len(files)):
            checked += 2
            for how, got in (("batch", batch[k]), ("single", single[k])):
                if got != want[k]:
                    if isinstance(want[k], str) and how == "batch":     # the reference raises for this file: a batch cannot mirror that
                        continue
                    info = dict(seed=seed, k=k, how=how, file=descs[k], az=az, md=md,
                                want=want[k] if isinstance(want[k], str) else len(want[k]),
                                got=got if isinstance(got, str) else len(got))
                    bad.append(info)
                    if verbose:
                        print("MISMATCH", json.dumps(info), flush=True)
    return checked, bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, nargs=2, default=[0, 40])
    a = ap.parse_args()
    t0 = time.time()
    res = sweep(*a.seeds)
    if res is None:
        print(json.dumps({"unavailable": "oracle/_ref/image_preprocessing.py is not in this snapshot"}))
        return
    checked, bad = res
    print(json.dumps({"seeds": a.seeds, "checked": checked, "mismatches": len(bad), "seconds": round(time.time() - t0, 1)}))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/sweep_files_vs_reference.json", "w") as f:
        json.dump({"seeds": a.seeds, "checked": checked, "mismatches": bad}, f)


if __name__ == "__main__":
    main()
