"""Randomised sweep of the device JPEG codec against Pillow (libjpeg-turbo) on a GPU box: random sizes, contents
(noise, gradients, text pages, flat, high-contrast stripes), qualities and optimize flags.
 * encode: the FILE the device writes must equal Pillow's byte for byte (image_preprocessing.py:312-347, 496-557);
 * decode: the raster the device decodes from Pillow's file must equal Pillow's decode (:57-75), for 4:2:0 and 4:4:4.

    python tools/sweep_jpeg_vs_pillow.py --seeds 0 150
"""
import argparse
import io
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from PIL import Image

import oracle as O
from ocr_system_b200 import ops


def content(rng, seed, h, w):
    kind = seed % 5
    yy, xx = np.mgrid[0:h, 0:w]
    if kind == 0:
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    if kind == 1:
        return np.stack([(xx * 3 + yy) % 256, (yy * 2) % 256, (xx + yy * 5) % 256], -1).astype(np.uint8)
    if kind == 2:
        return O.synth_page(h, w, seed)
    if kind == 3:
        return np.full((h, w, 3), int(rng.integers(0, 256)), np.uint8)
    a = np.where(((xx // int(rng.integers(1, 9))) + (yy // int(rng.integers(1, 9)))) % 2 == 0, 0, 255).astype(np.uint8)
    return np.stack([a, 255 - a, a], -1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, nargs=2, default=[0, 100])
    a = ap.parse_args()
    bad, checked, t0 = [], 0, time.time()
    for seed in range(*a.seeds):
        rng = np.random.default_rng(seed)
        h, w = int(rng.integers(1, 700)), int(rng.integers(1, 900))
        if seed % 11 == 0:
            h, w = int(rng.integers(900, 2100)), int(rng.integers(900, 1500))
        img = content(rng, seed, h, w)
        q = int(rng.choice([30, 40, 55, 65, 75, 85, 95, 100, 1]))
        opt = bool(rng.integers(0, 2))
        try:
            b = io.BytesIO()
            Image.fromarray(img).save(b, format="JPEG", quality=q, optimize=opt)
            ref = b.getvalue()
        except OSError:            # Pillow's own output buffer is too small for incompressible data without optimize:
            ref = None             # the reference raises here too; nothing to compare
        if ref is not None:
            got = ops.jpeg_encode(torch.from_numpy(img[None]).cuda(), q, opt)[0]
            checked += 1
            if got != ref:
                bad.append((seed, "encode", h, w, q, opt))
                print("MISMATCH encode", seed, h, w, q, opt, len(ref), len(got), flush=True)
        for sub in (2, 0):      # 4:2:0 (Pillow's default below q 100... keep explicit) and 4:4:4
            try:
                b = io.BytesIO()
                Image.fromarray(img).save(b, format="JPEG", quality=q, optimize=opt, subsampling=sub)
                data = b.getvalue()
            except OSError:
                continue
            want = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
            try:
                dec = ops.jpeg_decode([data]).cpu().numpy()[0]
                ok = dec.shape == want.shape and np.array_equal(dec, want)
            except Exception as e:  # noqa: BLE001
                ok = False
                print("decode raised", repr(e)[:200], flush=True)
            checked += 1
            if not ok:
                bad.append((seed, "decode", h, w, q, opt, sub))
                print("MISMATCH decode", seed, h, w, q, opt, "subsampling", sub, flush=True)
    print(json.dumps({"seeds": a.seeds, "checked": checked, "mismatches": len(bad), "seconds": round(time.time() - t0, 1)}))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/sweep_jpeg_vs_pillow.json", "w") as f:
        json.dump({"seeds": a.seeds, "checked": checked, "mismatches": bad}, f)


if __name__ == "__main__":
    main()
