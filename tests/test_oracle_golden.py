"""Pins the CPU oracle (oracle/lumina_oracle.c) to the REAL reference: every case in
tests/golden/reference_golden.json was produced by the unmodified reference module
(tests/golden/make_golden.py); the oracle must reproduce each output byte for byte."""
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_golden.json")) as f:
    GOLD = json.load(f)


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


_pages = {}


def page(O, c):
    k = (c["h"], c["w"], c["seed"])
    if k not in _pages:
        _pages[k] = O.synth_page(*k)
    return _pages[k]


def _resized(O, c):
    rgb = page(O, c)
    tw, th = O.target_size(c["w"], c["h"], c["max_dim"])
    return rgb if (tw, th) == (c["w"], c["h"]) else O.resize_lanczos(rgb, tw, th)


SMALL = [c for c in GOLD["cases"] if c["page"] != "a4_300dpi"]
FULL = [c for c in GOLD["cases"] if c["page"] == "a4_300dpi"]


def _check(O, c):
    st = c["stage"]
    rgb = page(O, c)
    if st == "resize_if_needed":
        out = _resized(O, c)
        assert list(out.shape[1::-1]) == c["size"]
        assert sha(out) == c["sha"]
    elif st == "resize_if_needed_L":
        g = O.gray_pil(rgb)
        tw, th = O.target_size(c["w"], c["h"], c["max_dim"])
        assert sha(O.resize_lanczos(g, tw, th)) == c["sha"]
    elif st in ("deskew", "deskew_full"):
        r = _resized(O, c)
        img, angle, lines = O.deskew(r)
        assert angle == c["angle"]
        assert sha(img) == c["sha"]
        if st == "deskew":
            g = O.gray_cv(r)
            e = O.canny(g, 50, 150)
            assert sha(g) == c["gray_cv_sha"] and sha(e) == c["canny_sha"]
            assert len(lines) == c["n_lines"] and sha(lines.astype(np.int32)) == c["lines_sha"]
    elif st == "azure_chain_before_jpeg":
        img, _, _ = O.deskew(_resized(O, c))
        assert sha(O.sharpness(O.contrast(img, 1.2), 1.1)) == c["sha"]
    elif st == "optimize_for_ocr":
        assert sha(O.sharpness(O.contrast(_resized(O, c), 1.2), 1.1)) == c["sha"]
    elif st == "adaptive_binarize":
        assert sha(O.adaptive_gauss11(O.gray_pil(_resized(O, c)), 2)) == c["sha"]
    elif st == "convert_to_grayscale":
        assert sha(O.gray_pil(rgb)) == c["sha"]
    elif st == "enhance_contrast":
        assert sha(O.contrast(rgb, c["factor"])) == c["sha"]
    elif st == "enhance_contrast_L":
        assert sha(O.contrast(O.gray_pil(rgb), c["factor"])) == c["sha"]
    elif st == "enhance_sharpness":
        assert sha(O.sharpness(rgb, c["factor"])) == c["sha"]
    elif st == "denoise":
        assert sha(O.median3(rgb)) == c["sha"]
    elif st == "binarize":
        assert c["mode"] == "1"
        assert sha(O.threshold(O.gray_pil(rgb), 128)) == c["sha"]
    elif st == "auto_orient":
        out = O.exif_transpose(rgb, c["orientation"])
        assert list(out.shape[1::-1]) == c["size"] and sha(out) == c["sha"]
    else:
        raise AssertionError(f"unknown golden stage {st}")


@pytest.mark.parametrize("idx", range(len(SMALL)))
def test_oracle_matches_reference_small(oracle, idx):
    _check(oracle, SMALL[idx])


@pytest.mark.parametrize("idx", range(len(FULL)))
def test_oracle_matches_reference_a4(oracle, idx):
    _check(oracle, FULL[idx])


def test_golden_covers_every_hot_path_stage():
    stages = {c["stage"] for c in GOLD["cases"]}
    need = {"resize_if_needed", "convert_to_grayscale", "enhance_contrast", "enhance_sharpness", "denoise", "binarize",
            "adaptive_binarize", "deskew", "auto_orient", "optimize_for_ocr", "azure_chain_before_jpeg"}
    assert need <= stages
    assert GOLD["versions"]["Pillow"] and GOLD["versions"]["opencv"]
